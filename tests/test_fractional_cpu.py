"""Fractional view-groups (a group's K views spread over several ranks; train.py:185-209, sac.py:198-216,243-245).
CPU only: the oracle's restatement against golden vectors the REAL reference produced on two gloo ranks
(tests/golden/make_golden_fractional.py), and the host-side index logic on a world_size-2 gloo group."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
K, HW, WORLD = 4, (96, 96), 2


def load_golden():
    return np.load(os.path.join(HERE, "golden", "sac_fractional_w2.npz"))


def test_oracle_fractional_refine_matches_reference_on_two_ranks():
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    g = load_golden()
    cfg = synth.ModelCfg()
    batch = synth.make_target_batch(1, K, HW, seed=3)
    per = K // WORLD
    parts = [[t[r * per:(r + 1) * per] for t in batch] for r in range(WORLD)]
    logits = [torch.from_numpy(g["r%d_teacher_logits" % r]) for r in range(WORLD)]
    ign = [(p[1] == -1) for p in parts]
    rc0 = [torch.full((19,), cfg.THRESHOLD_BETA) for _ in range(WORLD)]
    res = O.refine_fractional(logits, HW, K, [p[3] for p in parts], [p[4] for p in parts], ign, rc0, cfg, training=True)
    for r, (refined, rc) in enumerate(res):
        assert torch.allclose(rc, torch.from_numpy(g["r%d_running_conf" % r]), rtol=1e-5, atol=1e-8)
        ref_sub = torch.from_numpy(g["r%d_teacher_refined_sub" % r])
        assert (refined[:, :, ::3, ::3] - ref_sub).abs().max() < 2e-6
        labels, conf, _, _ = O.pseudo_labels_probs(refined, ign[r], rc, cfg, cfg.CONF_DISCOUNT)
        assert (conf - torch.from_numpy(g["r%d_teacher_conf" % r])).abs().max() < 2e-6
        gold = torch.from_numpy(g["r%d_teacher_labels" % r].astype(np.int64))
        amb = torch.from_numpy(g["r%d_ambiguous" % r])
        assert int(((labels != gold) & ~amb).sum()) == 0
        assert (labels != 255).float().mean() > 0.1


def test_fractional_pool_equals_whole_group_pool():
    """splitting a group over ranks must not change the result beyond fp32 summation order"""
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    cfg = synth.ModelCfg()
    torch.manual_seed(5)
    batch = synth.make_target_batch(1, K, (48, 48), seed=4)
    lg = torch.randn(K, 19, 7, 7) * 4
    ign = batch[1] == -1
    rc = torch.full((19,), cfg.THRESHOLD_BETA)
    whole, _, _ = O.refine(lg, (48, 48), K, batch[3], batch[4], ign, rc, cfg, training=False)
    parts = O.refine_fractional([lg[:2], lg[2:]], (48, 48), K, [batch[3][:2], batch[3][2:]], [batch[4][:2], batch[4][2:]],
                                [ign[:2], ign[2:]], [rc, rc], cfg, training=False)
    got = torch.cat([p[0] for p in parts], 0)
    assert (got - whole).abs().max() < 1e-6


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from da_sac_b200.trainer import prep_batch, fractional_subgroup
    N, L = 1, 4                                            # 1 group x 4 views on 2 ranks -> 2 views per rank
    mine = torch.arange(L, dtype=torch.float32).view(1, L, 1) + 100.0 * rank      # this rank's loader batch [B=1,T,...]
    out = prep_batch(mine, N, L, world, rank)
    # train.py:196-209: both ranks take their slice of RANK 0's batch (index0 = rank*2 // 4 = 0)
    ok = torch.equal(out.flatten(), torch.tensor([2.0 * rank, 2.0 * rank + 1]))
    ok = ok and fractional_subgroup(rank, 2, 4) == (0, 2)
    # whole groups per rank: plain flatten, no communication
    whole = prep_batch(mine.repeat(2, 1, 1), 4, L, world, rank)
    ok = ok and whole.shape == (8, 1)
    # partial sums exchange = sum over the sub-group
    part = torch.full((10,), float(rank + 1))
    dist.all_reduce(part, group=dist.new_group([0, 1]))
    ok = ok and torch.equal(part, torch.full((10,), 3.0))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_prep_batch_and_subgroups_on_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_fractional_subgroup_index_math():
    from da_sac_b200.trainer import fractional_subgroup
    # 2 groups x 4 views on 4 GPUs (the reference's default recipe): ranks 0,1 share group 0; ranks 2,3 share group 1
    assert [fractional_subgroup(r, 2, 4) for r in range(4)] == [(0, 2), (0, 2), (2, 2), (2, 2)]
    # 1 view per rank, groups of 4 on 8 ranks
    assert [fractional_subgroup(r, 1, 4)[0] for r in range(8)] == [0, 0, 0, 0, 4, 4, 4, 4]
    assert fractional_subgroup(3, 4, 4) == (3, 1)          # whole group on the rank: no exchange


def _worker4(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    from da_sac_b200.trainer import prep_batch
    net = get_model(synth.ModelCfg(), rank, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    N, L = 2, 4                                            # the reference's default recipe: 2 groups x 4 views on 4 GPUs
    # every rank's loader hands it ONE group ([1, T, ...], datasets/__init__.py:64); ranks 0,1 must end up with halves of
    # rank 0's group, ranks 2,3 with halves of rank 1's (train.py:196-209)
    mine = torch.arange(L, dtype=torch.float32).view(1, L, 1) + 100.0 * rank
    out = prep_batch(mine, N, L, world, rank)
    src, half = (rank * 2) // L, (rank * 2) % L
    ok = torch.equal(out.flatten(), torch.tensor([100.0 * src + half, 100.0 * src + half + 1]))
    # the product's exchange of the reference-frame partial sums: a sum over the ranks that share the group, nobody else
    pooled = torch.full((7,), float(10 ** rank))
    net._exchange_partial_sums(pooled, 2, L)
    want = {0: 11.0, 1: 11.0, 2: 1100.0, 3: 1100.0}[rank]
    ok = ok and torch.equal(pooled, torch.full((7,), want))
    net._exchange_partial_sums(pooled, 2, L)               # second call re-uses the cached sub-groups
    ok = ok and torch.equal(pooled, torch.full((7,), 2 * want))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_default_recipe_two_groups_of_four_views_on_four_ranks():
    """configs/deeplabv2_resnet101_train.yaml:16-18 on 4 GPUs: `_prep_batch` slicing and `SAC._exchange_partial_sums` with two
    sub-groups {0,1} and {2,3} (every rank creates every sub-group, in the same order: new_group is collective)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker4, args=(r, 4, port, q)) for r in range(4)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res == [(0, True), (1, True), (2, True), (3, True)]
