"""GPU parity of the ABN baseline mode (cfg.MODEL.BASELINE = True: training-mode BN, engine_abn.py + csrc/sacb_bn.cu).

(1) the BN kernels through the C ABI against torch.nn.functional.batch_norm in fp64 (forward, running statistics, backward);
(2) one iteration of the reference's BASELINE recipe -- source step with SGD, no-grad target pass, eval forward -- against the
    golden vectors produced by the REAL reference (tests/golden/make_golden_abn.py).

First run on a B200 in round 2 (7 passed, profiles/r2a_test_abn_gpu.log); the oracle for this mode is pinned on the CPU
(tests/test_abn_cpu.py).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import bn_kernel_checks as K
from bn_kernel_checks import join, rel, split  # noqa: F401

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS, MOM = 1e-5, 0.1


@pytest.mark.parametrize("M,Cn,with_res", K.CASES)
def test_bn_kernels_match_torch_fp64(M, Cn, with_res):
    from da_sac_b200 import lib as L
    K.bn_kernels_match_torch_fp64(K.Api(L.lib(), L.stream(), L.ptr, torch.cuda.synchronize, "cuda"), M, Cn, with_res)


def test_bn_moments_survive_large_mean():
    from da_sac_b200 import lib as L
    K.bn_moments_survive_large_mean(K.Api(L.lib(), L.stream(), L.ptr, torch.cuda.synchronize, "cuda"))


ARCHS = {   # arch -> (golden, state_dict factory name, cfg name, (N_SRC, N_TGT, HW), nbt key, probe weight, min launches)
    "resnet101": ("abn_resnet101_tiny.npz", "make_backbone_params", 123, "ModelCfg", (4, 3, (129, 129)),
                  "model.layer3.5.bn2.num_batches_tracked", "model.layer3.5.conv2.weight", 800),
    "vgg16": ("abn_vgg16_tiny.npz", "make_vgg16_params", 321, "ModelCfgVGG16", (3, 2, (96, 96)),
              "features.18.num_batches_tracked", "features.17.weight", 150),
    "fcn": ("abn_fcn8s_tiny.npz", "make_fcn_params", 213, "ModelCfgFCN", (3, 2, (96, 96)),
            "vgg_head.1.num_batches_tracked", "block2.27.weight", 150),
}


@pytest.mark.parametrize("arch", ["resnet101", "vgg16", "fcn"])
def test_abn_iteration_matches_reference_golden(arch):
    """source step (training BN + SGD) -> no-grad target pass (statistics only) -> eval forward, vs the real reference"""
    from da_sac_b200 import lib as L, synth
    from da_sac_b200.models import get_model
    fname, make_sd, seed, cfg_name, (N_SRC, N_TGT, HW), nbt_key, probe, min_launches = ARCHS[arch]
    g = np.load(os.path.join(ROOT, "tests", "golden", fname), allow_pickle=False)

    class Cfg(getattr(synth, cfg_name)):
        BASELINE = True

    cfg = Cfg()
    extra = {"drop_rate": 0.0} if arch == "fcn" else {}
    net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"), **extra)
    assert type(net).__name__ == "SAC_Baseline"
    net.backbone.load_state_dict(getattr(synth, make_sd)(seed=seed))
    net.cuda().train()
    optim = torch.optim.SGD(net.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    xs, ys = [t.cuda() for t in synth.make_source_batch(N_SRC, HW, seed=0)]
    xt, yt = [t.cuda() for t in synth.make_source_batch(N_TGT, HW, seed=1)]
    n0 = L.launch_count()
    # ---- source step (train.py:119-138)
    losses, outs = net(xs, ys)
    optim.zero_grad()
    losses["loss_ce"].mean().backward()
    torch.cuda.synchronize()
    assert L.launch_count() - n0 > min_launches, "the training-BN CUDA path did not run"
    l2, mx = rel(outs["logits"].detach(), g["src_logits"])
    print(arch, "ABN source logits rel-L2 %.2e max %.2e" % (l2, mx))
    assert l2 < 1e-3 and mx < 1e-3
    ref_loss = float(g["src_loss_ce"].reshape(-1)[0])
    assert abs(float(losses["loss_ce"].detach()) - ref_loss) < 2e-3 * ref_loss
    params = dict(net.backbone.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    mine = np.array([params[n].grad.double().norm().item() for n in names])
    gn = g["src_grad_norms"]
    big = gn > 1e-6 * gn.max()          # conv biases in front of a training-mode BN: zero gradient up to rounding noise
    relerr = np.abs(mine - gn)[big] / gn[big]
    print(arch, "ABN grad-norm max rel err %.2e" % relerr.max())
    assert relerr.max() < 2e-2 and np.all(mine[~big] <= 1e-5 * gn.max())
    for key in g.files:
        if key.startswith("src_grad::") and np.abs(g[key]).max() > 1e-6 * gn.max():
            n = key.split("::")[1]
            gr = params[n].grad
            gr = gr.flatten()[:60000] if gr.numel() > 60000 else gr
            e = rel(gr.reshape(g[key].shape), g[key])[0]
            print("   grad", n, "rel-L2 %.2e" % e)
            # early layers: 2-3e-2, set by ReLU / max-pool decisions that flip under the 2^-17 plane rounding, not by the arithmetic
            # (the reference's own fp32 gradients are 2-4e-3 from an fp64 run; the same network with a different fp32 summation
            # order moves this number between 2.7e-2 and 3.0e-2 on features.18.bias: formula model vs tcgen05 emulation)
            assert e < 5e-2, key
    optim.step()
    sd = net.backbone.state_dict()
    stat_names = [str(k) for k in g["stat_names"]]
    stats = torch.cat([sd[k].reshape(-1) for k in stat_names])
    assert rel(stats, g["src_stats"])[1] < 1e-4
    assert int(sd[nbt_key]) == int(g["src_nbt"])
    for key in g.files:
        if key.startswith("src_post::"):
            e = rel(sd[key.split("::")[1]].flatten()[:60000], g[key])[1]
            print("   post-SGD", key.split("::")[1], "max-rel %.2e" % e)
            # the update is lr * grad and the early-layer gradients sit 2-3e-2 from the reference's (ReLU / max-pool decisions
            # flipped by the 2^-17 plane rounding; the reference's own fp32 gradients are 2-4e-3 from fp64 there)
            assert e < 5e-5, key
    # ---- ABN target pass (train.py:281-289): statistics only
    # Two yardsticks.  (a) The pinned oracle run on OUR post-step weights: same weights in, so this isolates the forward
    # schedule and holds the tight bar.  (b) The reference's golden: its weights took the reference's SGD step, ours took ours,
    # and the step's gradients differ by the plane-rounding noise discussed above; FCN-8s' 4096-wide head on 27 samples per
    # channel (lr x 10) turns that into 4e-3 on the next forward's logits, the other two backbones stay below 1e-3.
    from oracle import sac_oracle as O
    oracle_p = O.as_leaf_params({k: v.detach().cpu().clone() for k, v in sd.items()})
    w_before = sd[probe].detach().clone()
    with torch.no_grad():
        losses_t, outs_t = net(xt, yt)
    losses_o, outs_o = O.baseline_target_pass(oracle_p, xt.cpu(), yt.cpu())
    e_o, e_g = rel(outs_t["logits"], outs_o["logits"])[1], rel(outs_t["logits"], g["tgt_logits"])[1]
    print(arch, "ABN target-pass logits: vs oracle on the same weights %.2e, vs golden %.2e" % (e_o, e_g))
    golden_bar = {"fcn": 1e-2, "vgg16": 3e-3, "resnet101": 1e-3}[arch]
    assert e_o < 1e-3 and e_g < golden_bar
    assert abs(float(losses_t["loss_ce"]) - float(losses_o["loss_ce"])) < 1e-3 * float(losses_o["loss_ce"])
    ref_loss = float(g["tgt_loss_ce"].reshape(-1)[0])
    assert abs(float(losses_t["loss_ce"]) - ref_loss) < 2 * golden_bar * ref_loss
    sd = net.backbone.state_dict()
    stats = torch.cat([sd[k].reshape(-1) for k in stat_names])
    assert rel(stats, torch.cat([oracle_p[k].detach().reshape(-1) for k in stat_names]))[1] < 1e-4
    assert rel(stats, g["tgt_stats"])[1] < (2e-3 if arch == "fcn" else 1e-4)
    assert int(sd[nbt_key]) == int(g["tgt_nbt"])
    assert torch.equal(w_before, sd[probe])
    # ---- evaluation with the adapted statistics: the frozen-BN engine
    net.eval()
    logits_e, _ = net(xt)
    with torch.no_grad():
        logits_eo, _ = O.backbone_forward(oracle_p, xt.cpu())
    e_o, e_g = rel(logits_e, logits_eo)[1], rel(logits_e, g["eval_logits"])[1]
    print(arch, "ABN eval logits: vs oracle on the same weights %.2e, vs golden %.2e" % (e_o, e_g))
    assert e_o < 1e-3 and e_g < golden_bar
