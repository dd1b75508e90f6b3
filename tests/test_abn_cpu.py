"""ABN baseline mode (cfg.MODEL.BASELINE = True): pins the oracle's training-mode-BN restatement against golden vectors
produced by the real reference (tests/golden/make_golden_abn.py), and checks the host-side statistics exchange that
stands in for nn.SyncBatchNorm across ranks on two gloo processes.  CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from da_sac_b200 import synth
from oracle import sac_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCHS = {   # arch -> (golden file, state_dict factory, cfg, (N_SRC, N_TGT, HW), num_batches_tracked key, #BN layers)
    "resnet101": ("abn_resnet101_tiny.npz", lambda: synth.make_backbone_params(seed=123), synth.ModelCfg, (4, 3, (129, 129)),
                  "model.layer3.5.bn2.num_batches_tracked", 104),
    "vgg16": ("abn_vgg16_tiny.npz", lambda: synth.make_vgg16_params(seed=321), synth.ModelCfgVGG16, (3, 2, (96, 96)),
              "features.18.num_batches_tracked", 13),
    "fcn": ("abn_fcn8s_tiny.npz", lambda: synth.make_fcn_params(seed=213), synth.ModelCfgFCN, (3, 2, (96, 96)),
            "vgg_head.1.num_batches_tracked", 15),
}


def rel(a, b):
    a = torch.as_tensor(a).double(); b = torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _stats(params, names):
    return torch.cat([params[str(k)].detach().reshape(-1) for k in names])


@pytest.mark.parametrize("arch", sorted(ARCHS))
def test_oracle_abn_source_step_target_pass_and_eval_match_reference(arch):
    fname, make_sd, Cfg, (N_SRC, N_TGT, HW), nbt_key, n_bn = ARCHS[arch]
    g = np.load(os.path.join(ROOT, "tests", "golden", fname), allow_pickle=False)
    torch.set_num_threads(8)
    cfg = Cfg()
    student = O.as_leaf_params(make_sd())
    optim = torch.optim.SGD(O.parameter_groups(student, cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    xs, ys = synth.make_source_batch(N_SRC, HW, seed=0)
    xt, yt = synth.make_source_batch(N_TGT, HW, seed=1)
    # ---- source step: batch statistics, loss_ce backward, SGD, running statistics moved by momentum 0.1
    losses, outs = O.baseline_source_step(student, xs, ys, optim)
    l2, mx = rel(outs["logits"].detach(), g["src_logits"])
    assert l2 < 1e-4 and mx < 1e-4, (l2, mx)
    assert abs(float(losses["loss_ce"].detach()) - float(g["src_loss_ce"].reshape(-1)[0])) < 1e-5
    names = [str(n) for n in g["grad_names"]]
    mine = np.array([student[n].grad.double().norm().item() for n in names])
    gn = g["src_grad_norms"]
    # a conv bias in front of a training-mode BN has a mathematically zero gradient: both sides hold rounding noise there
    # (VGG convs; ResNet convs have no bias), so those entries are compared absolutely
    big = gn > 1e-6 * gn.max()
    assert np.all(np.abs(mine - gn)[big] <= 2e-3 * gn[big]), np.max(np.abs(mine - gn)[big] / gn[big])
    assert np.all(mine[~big] <= 1e-5 * gn.max())
    for key in g.files:
        if key.startswith("src_grad::"):
            n = key.split("::")[1]
            gr = student[n].grad.flatten()[:60000] if student[n].grad.numel() > 60000 else student[n].grad
            if np.abs(g[key]).max() > 1e-6 * gn.max():
                assert rel(gr.reshape(g[key].shape), g[key])[0] < 2e-3, key
        if key.startswith("src_post::"):
            n = key.split("::")[1]
            assert rel(student[n].detach().flatten()[:60000], g[key])[1] < 1e-6, key
    stat_names = [str(k) for k in g["stat_names"]]
    assert len(stat_names) == 2 * n_bn
    l2, mx = rel(_stats(student, stat_names), g["src_stats"])
    assert l2 < 1e-5 and mx < 1e-5, (l2, mx)
    assert int(student[nbt_key]) == int(g["src_nbt"]) == 1
    # ---- ABN target pass: no gradient, only the running statistics change
    before = {k: v.detach().clone() for k, v in student.items() if k.endswith(".weight") or k.endswith(".bias")}
    losses_t, outs_t = O.baseline_target_pass(student, xt, yt)
    assert rel(outs_t["logits"], g["tgt_logits"])[1] < 1e-4
    assert abs(float(losses_t["loss_ce"]) - float(g["tgt_loss_ce"].reshape(-1)[0])) < 1e-5
    assert rel(_stats(student, stat_names), g["tgt_stats"])[1] < 1e-5
    assert int(student[nbt_key]) == int(g["tgt_nbt"]) == 2
    assert all(torch.equal(before[k], student[k].detach()) for k in before)
    # ---- evaluation with the adapted statistics (frozen-BN path of the same oracle)
    with torch.no_grad():
        logits_e, _ = O.backbone_forward(student, xt)
    assert rel(logits_e, g["eval_logits"])[1] < 1e-4


def test_training_bn_differs_from_frozen_bn():
    """guards against a silent fallback: on this fixture batch statistics and running statistics give different logits"""
    student = synth.make_backbone_params(seed=123)
    xs, ys = synth.make_source_batch(2, (65, 65), seed=3)
    with torch.no_grad():
        (lt, _), _, _ = O.baseline_forward(student, xs)
        le, _ = O.backbone_forward(student, xs)
    assert rel(lt, le)[0] > 0.05


# ---------------------------------------------------------------- SyncBatchNorm across ranks = BN over the global batch
def _worker(rank, world, port, out):
    import torch.distributed as dist
    from da_sac_b200 import trainer
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7)
    z = torch.randn(6, 5, 4, 4) * 2 + 1                    # the same global batch on both ranks; rank r owns samples r::world
    mine = z[rank::world]
    sums = torch.stack([mine.double().sum(dim=(0, 2, 3)), (mine.double() ** 2).sum(dim=(0, 2, 3))])
    count = torch.tensor([mine.numel() // mine.shape[1]], dtype=torch.float64)
    tot, n = trainer.sync_bn_sums(sums, count)
    mean = tot[0] / n
    var = tot[1] / n - mean * mean
    ref_mean = z.double().mean(dim=(0, 2, 3)); ref_var = z.double().var(dim=(0, 2, 3), unbiased=False)
    out[rank] = (float((mean - ref_mean).abs().max()), float((var - ref_var).abs().max()), float(n))
    dist.destroy_process_group()


def test_sync_bn_sums_world2_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29631, out), nprocs=2, join=True)
    for r in (0, 1):
        dm, dv, n = out[r]
        assert dm < 1e-12 and dv < 1e-12 and n == 6 * 16
