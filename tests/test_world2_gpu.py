"""Two real GPUs, NCCL: the drop-in under the reference's own wrapper, and the fractional-group path on real ranks.

(1) ``check_dropin_under_ddp_and_torch_sgd``: what /root/reference/train.py does with the model, line by line, with
    ``da_sac_b200.models`` in place of the reference's ``models``: ``get_model`` -> ``base_trainer.get_optim``
    (torch.optim.SGD over ``net.parameter_groups``, base_trainer.py:61-66) -> ``net.cuda(gpu)`` ->
    ``DistributedDataParallel(net, device_ids=[gpu])`` (train.py:100-104) -> two ``_step_target`` iterations
    (train.py:211-250) including the IN-PLACE all-reduce / division of every loss tensor (train.py:243-246).  Every rank holds
    one view-group.  Expected values come from the CPU oracle run per rank: DDP averages the per-rank gradients, so the
    parameters after step 0 are SGD(mean of the two oracle gradients) and step 1 runs on them; DDP broadcasts rank 0's buffers
    before every forward, which makes rank 0's ``running_conf`` authoritative.
(2) ``check_fractional_group_on_two_ranks``: the reference's default recipe gives a GPU only part of a view-group
    (train.py:185-209, sac.py:198-216).  1 group x K=4 views on two ranks, two views each, the sub-group all-reduce of
    ``SAC._exchange_partial_sums`` on NCCL, against the golden vectors the REAL reference produced on two gloo ranks
    (tests/golden/make_golden_fractional.py).

Run on a 2-GPU box: ``gpurun --gpus 2 -- python -m pytest tests/test_world2_gpu.py -q``; log kept in profiles/.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")]

HERE = os.path.dirname(os.path.abspath(__file__))
WORLD = 2


def rel(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _spawn(fn, tag):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (17 if tag == "frac" else 0)
    procs = [ctx.Process(target=_guard, args=(fn, r, port, q)) for r in range(WORLD)]
    for p in procs: p.start()
    res = sorted((q.get(timeout=900) for _ in procs), key=lambda t: t[0])
    for p in procs: p.join(timeout=120)
    for rank, out, msg in res:
        assert msg == "", "rank %d: %s" % (rank, msg)
    return [out for _, out, _ in res]


def _plain(o):
    """tensors -> numpy: the result travels through a multiprocessing queue by value (torch would share file descriptors of a
    process that may already have exited)"""
    if isinstance(o, torch.Tensor):
        return o.detach().cpu().numpy()
    if isinstance(o, dict):
        return {k: _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(_plain(v) for v in o)
    return o


def _guard(fn, rank, port, q):
    import traceback
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    try:
        q.put((rank, _plain(fn(rank, dev)), ""))
    except Exception:          # surface the failure instead of hanging the parent
        q.put((rank, None, traceback.format_exc()[-3000:]))
    try:
        dist.destroy_process_group()
    except Exception:
        pass


# ------------------------------------------------------------------------------------------------ (1) DDP + torch.optim.SGD
N_GROUPS, K, HW = 2, 2, (128, 128)


def _ddp_rank(rank, dev):
    import torch.distributed as dist
    from da_sac_b200 import lib as L, synth
    from da_sac_b200.models import get_model
    cfg = synth.ModelCfg()
    net = get_model(cfg, rank, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))   # train.py:88-89
    net.backbone.load_state_dict(synth.make_backbone_params(seed=123))
    # base_trainer.get_optim (base_trainer.py:61-66), called on the un-wrapped net before it moves to the GPU (train.py:92)
    optim = torch.optim.SGD(net.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY), lr=cfg.LR, momentum=cfg.MOMENTUM,
                            nesterov=False, weight_decay=cfg.WEIGHT_DECAY)
    net.cuda(rank)                                                                          # train.py:103
    ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[rank])                 # train.py:104
    ddp.train()                                                                             # train.py:263
    batch = synth.make_target_batch(N_GROUPS, K, HW, seed=0)
    mine = [t[rank * K:(rank + 1) * K] for t in batch]          # one whole view-group per rank: _prep_batch's early-out (train.py:186-187)
    n0 = L.launch_count()
    out = {}
    for step in (0, 1):
        frames1, frames_gt, frames2, affine, affine_inv = [t.clone().cuda(rank, non_blocking=True) for t in mine]
        losses, logits = ddp(frames1, frames_gt, frames2, affine, affine_inv, use_teacher=True,
                             update_teacher=(step == 0), T=K)                               # train.py:219-222, :294
        optim.zero_grad()                                                                   # train.py:227-228
        loss_target = cfg.LR_TARGET * losses["self_ce"].mean()                              # train.py:231
        loss_target.backward()
        optim.step()                                                                        # train.py:233
        for key, val in losses.items():                                                     # train.py:243-246
            dist.all_reduce(val)
            val /= WORLD
            losses[key] = losses[key].item()
        logits["mask_gt"] = frames_gt                                                       # train.py:249
        torch.cuda.synchronize()
        pre = "s%d_" % step
        out[pre + "losses"] = dict(losses)
        out[pre + "logits"] = logits["logits"].detach().cpu()
        out[pre + "labels"] = logits["teacher_labels"].cpu()
        out[pre + "running_conf"] = net.running_conf.detach().cpu().clone()
        out[pre + "mask_gt_ok"] = bool((frames_gt != -1).all())                             # in-place -1 -> 255 on the caller's tensor
        for k in ("teacher_aligned", "frames_aligned", "teacher_refined", "teacher_init", "teacher_conf", "logits_up"):
            assert k in logits, k                                                           # the reference's net_outs keys (sac.py:293-296,362-371)
        if step == 0:
            ta, fa = logits["teacher_aligned"], logits["frames_aligned"]
            out["aligned_shapes"] = (tuple(ta.shape), tuple(fa.shape))
    out["launches"] = L.launch_count() - n0
    out["params"] = {k: v.detach().cpu().clone() for k, v in net.backbone.named_parameters() if k in PICKS}
    flat = net.backbone._flat.buf
    gathered = [torch.empty_like(flat) for _ in range(WORLD)]
    dist.all_gather(gathered, flat)
    out["replicas_equal"] = all(torch.equal(gathered[0], t) for t in gathered[1:])
    return out


PICKS = ("model.conv1.weight", "model.bn1.weight", "model.layer1.0.conv1.weight", "model.layer2.0.downsample.0.weight",
         "model.layer3.5.bn2.bias", "model.layer3.5.conv2.weight", "model.layer4.2.conv3.weight",
         "model.layer5.conv2d_list.1.bias", "model.layer5.conv2d_list.3.weight")


def _oracle_expectation():
    """two per-rank oracle steps -> mean gradient -> SGD -> step-1 forward per rank on the updated weights"""
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    cfg = synth.ModelCfg()
    sd = synth.make_backbone_params(seed=123)
    batch = synth.make_target_batch(N_GROUPS, K, HW, seed=0)
    ranks, exp = [], {}
    for r in range(WORLD):
        student = O.as_leaf_params(sd)
        teacher = {k: v.detach().clone() for k, v in student.items()}
        rc = torch.full((19,), cfg.THRESHOLD_BETA)
        mine = tuple(t[r * K:(r + 1) * K].clone() for t in batch)
        losses, outs, rc = O.sac_target_step(student, teacher, rc, mine, K, cfg, optim=None)
        ranks.append(dict(student=student, teacher=teacher, rc=rc, losses=losses, outs=outs, batch=mine))
    exp["s0_losses"] = {k: float(sum(float(r["losses"][k]) for r in ranks) / WORLD) for k in ("self_ce", "loss_ce", "teacher_diff")}
    # DDP: every rank steps with the mean gradient
    master = ranks[0]["student"]
    for k, p in master.items():
        if p.grad is not None:
            p.grad = sum(r["student"][k].grad for r in ranks) / WORLD
    optim = torch.optim.SGD(O.parameter_groups(master, cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    optim.step()
    exp["params"] = {k: master[k].detach().clone() for k in PICKS}
    exp["update"] = {k: (master[k].detach() - sd[k]) for k in PICKS}
    rc0 = ranks[0]["rc"]                                        # DDP broadcasts rank 0's buffers before the next forward
    for r in range(WORLD):
        for k, p in master.items():
            p.grad = None
        with torch.no_grad():
            losses, outs, rc = O.sac_target_forward(master, ranks[0]["teacher"], rc0.clone(), tuple(t.clone() for t in ranks[r]["batch"]), K, cfg, True)
        exp["r%d_s1_logits" % r] = outs["logits"].detach()
        exp["r%d_s1_labels" % r] = outs["teacher_labels"]
        exp["r%d_s1_self_ce" % r] = float(losses["self_ce"])
        exp["r%d_s0_logits" % r] = ranks[r]["outs"]["logits"].detach()
        exp["r%d_s0_labels" % r] = ranks[r]["outs"]["teacher_labels"]
    exp["s1_self_ce"] = sum(exp["r%d_s1_self_ce" % r] for r in range(WORLD)) / WORLD
    exp["sd"] = sd
    return exp


def check_dropin_under_ddp_and_torch_sgd():
    res = _spawn(_ddp_rank, "ddp")
    exp = _oracle_expectation()
    for r, out in enumerate(res):
        assert out["launches"] > 600, "CUDA path did not run on rank %d" % r
        assert out["replicas_equal"], "replicas diverged under DDP + torch.optim.SGD"
        assert out["s0_mask_gt_ok"] and out["s1_mask_gt_ok"]
        assert out["aligned_shapes"] == ((K, 19) + HW, (K, 3) + HW)
        for step in (0, 1):
            l2, mx = rel(out["s%d_logits" % step], exp["r%d_s%d_logits" % (r, step)])
            agree = (torch.as_tensor(out["s%d_labels" % step]) == exp["r%d_s%d_labels" % (r, step)]).float().mean().item()
            print("rank %d step %d: logits rel-L2 %.2e max %.2e, label agreement %.6f" % (r, step, l2, mx, agree))
            assert l2 < 1e-3 and mx < 1e-3, (r, step, l2, mx)
            assert agree > 0.999, (r, step, agree)
        # the all-reduced, in-place divided losses (identical on both ranks) = mean of the per-rank oracle losses
        for k, g in exp["s0_losses"].items():
            v = out["s0_losses"][k]
            assert abs(v - g) <= 2e-3 * max(abs(g), 1e-3), (k, v, g)
        v, g = out["s1_losses"]["self_ce"], exp["s1_self_ce"]
        print("rank %d: step-1 self_ce %.6f vs oracle %.6f" % (r, v, g))
        assert abs(v - g) <= 5e-3 * max(abs(g), 1e-3), (v, g)
        # parameters after TWO optimiser steps are not in the expectation; compare after step 0 through the update of step 0:
        # step 1's update is small against the tolerance only for the BN / bias tensors, so the check uses what DDP + SGD left
        # after step 0 indirectly -- the step-1 logits above -- and directly the replicas' equality.
    # rank 0 == rank 1 bit for bit was asserted; the per-tensor update of step 0 is checked in the single-GPU golden test
    # (tests/test_step_gpu.py) and, across ranks, by the fused-exchange test (tests/test_p2p_gpu.py).


# ------------------------------------------------------------------------------------------------ (2) fractional groups
FK, FHW = 4, (96, 96)


def _frac_rank(rank, dev):
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    from da_sac_b200.trainer import prep_batch
    cfg = synth.ModelCfg()
    net = get_model(cfg, rank, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    net.backbone.load_state_dict(synth.make_backbone_params(seed=123))
    net.cuda(rank).train()
    batch = synth.make_target_batch(1, FK, FHW, seed=3)
    # what the reference's loaders hand every rank of a 1 x 4 recipe on 2 GPUs is the SAME [1, T, ...] batch (batch_target =
    # max(1, NUM_GROUPS // ngpus) = 1, datasets/__init__.py:64); _prep_batch all-gathers and keeps this rank's 2 views
    loader = [t.view(1, FK, *t.shape[1:]) for t in batch]
    x, y, x2, A, Ai = [prep_batch(t, 1, FK, WORLD, rank, device=dev) for t in loader]      # train.py:157-209 on NCCL
    assert x.shape[0] == FK // WORLD
    losses, outs = net(x, y, x2, A, Ai, use_teacher=True, update_teacher=True, T=FK)
    refined = outs["teacher_refined"]                                # lazy; must not need another collective
    torch.cuda.synchronize()
    return dict(logits=outs["logits"].detach().cpu(), labels=outs["teacher_labels"].cpu().to(torch.uint8),
                conf=outs["teacher_conf"].cpu(), rc=outs["running_conf"].detach().cpu().clone(),
                self_ce=float(losses["self_ce"]), refined_sub=refined[:, :, ::3, ::3].cpu())


def check_fractional_group_on_two_ranks():
    g = np.load(os.path.join(HERE, "golden", "sac_fractional_w2.npz"))
    res = _spawn(_frac_rank, "frac")
    for r, out in enumerate(res):
        l2, mx = rel(out["logits"], g["r%d_logits" % r])
        assert l2 < 1e-3 and mx < 1e-3, (r, l2, mx)
        # running_conf is a mean of softmax outputs of logits that agree to ~1e-5: same bar as the single-GPU golden test
        assert rel(out["rc"], g["r%d_running_conf" % r])[1] < 1e-4
        assert np.abs(out["conf"] - g["r%d_teacher_conf" % r]).max() < 1e-3
        agree = float((out["labels"] == g["r%d_teacher_labels" % r]).mean())
        amb = torch.from_numpy(g["r%d_ambiguous" % r])
        print("rank %d: logits rel-L2 %.2e, label agreement %.6f (ambiguous pixels %d)" % (r, l2, agree, int(amb.sum())))
        assert agree > 0.999
        assert np.abs(out["refined_sub"] - g["r%d_teacher_refined_sub" % r]).max() < 1e-3
        ce = float(g["r%d_self_ce" % r].reshape(-1)[0])
        assert abs(out["self_ce"] - ce) <= 5e-3 * max(abs(ce), 1e-3), (out["self_ce"], ce)


def test_world2_dropin_under_ddp_and_fractional_groups():
    """both checks of this file in one test (one skip on a single-GPU box instead of two); each prints its own numbers"""
    check_dropin_under_ddp_and_torch_sgd()
    check_fractional_group_on_two_ranks()
