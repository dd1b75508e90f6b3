"""Parity AT THE BENCHMARKED GEOMETRIES (BASELINE.json configs[1], [3], [4]) -- the golden fixtures are 128^2 / 256^2 crops, the
bench runs 24 x 512^2, 16 x 640^2 and 6 x 1024^2 per GPU (M = 101 400 output rows per layer, 397 pair tiles, split-K over
101 400 pixels).  Two yardsticks, both executed on the GPU box at test time (nothing is read from /root/reference):

(1) ``test_full_step_512_one_group_vs_cpu_oracle``: one whole training step (forward, pseudo labels, loss, backward) of
    1 view-group x K=3 x 512^2 against ``oracle/sac_oracle.py`` on the host CPU (~5 s): logits <= 1e-3, masks >= 0.999
    agreement (the only disagreeing pixels sit on a threshold), self_ce, every parameter's gradient norm.
(2) ``test_backbone_and_tail_at_bench_batch``: at the full per-GPU batch the CPU oracle would take minutes, so the SAME oracle
    functions run on the GPU in strict fp32 (``cudnn.allow_tf32 = False``: ATen / cuDNN FFMA kernels, none of this repo's code)
    and provide (a) the backbone logits of the whole batch and (b), from OUR teacher logits, the refined probabilities,
    thresholds and pseudo labels of the whole batch.  Bars: logits 1e-3 (max-norm and rel-L2); labels bit-exact outside the
    pixels whose confidence sits within 1e-5 of its threshold or whose two best classes tie within 1e-5 (audited as in
    tests/golden/make_golden.py).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).double(); b = torch.as_tensor(b).double().to(a.device)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _net(arch):
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    cfg = {"resnet101": synth.ModelCfg, "fcn": synth.ModelCfgFCN, "vgg16": synth.ModelCfgVGG16}[arch]()
    sd = {"resnet101": lambda: synth.make_backbone_params(seed=123), "fcn": lambda: synth.make_fcn_params(seed=213),
          "vgg16": lambda: synth.make_vgg16_params(seed=321)}[arch]()
    extra = {"drop_rate": 0.0} if arch == "fcn" else {}
    m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"), **extra)
    m.backbone.load_state_dict(sd)
    m.cuda().train()
    return m, cfg, sd


def test_full_step_512_one_group_vs_cpu_oracle():
    import os
    from da_sac_b200 import lib as L, synth
    from oracle import sac_oracle as O
    G, K, HW = 1, 3, (512, 512)
    m, cfg, sd = _net("resnet101")
    batch = synth.make_target_batch(G, K, HW, seed=0)
    n0 = L.launch_count()
    x, y, x2, A, Ai = [t.clone().cuda() for t in batch]
    losses, outs = m(x, y, x2, A, Ai, use_teacher=True, update_teacher=True, T=K)
    (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
    torch.cuda.synchronize()
    assert L.launch_count() - n0 > 300
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    student = O.as_leaf_params(sd)
    teacher = {k: v.detach().clone() for k, v in student.items()}
    rc = torch.full((19,), cfg.THRESHOLD_BETA)
    ref_losses, ref_outs, _ = O.sac_target_step(student, teacher, rc, batch, K, cfg, optim=None)
    l2, mx = rel(outs["logits"].detach().cpu(), ref_outs["logits"].detach())
    agree = (outs["teacher_labels"].cpu() == ref_outs["teacher_labels"]).float().mean().item()
    valid = (ref_outs["teacher_labels"] != 255).float().mean().item()
    print("512^2 x3: logits rel-L2 %.2e max %.2e, label agreement %.6f (valid fraction %.3f)" % (l2, mx, agree, valid))
    assert l2 < 1e-3 and mx < 1e-3
    assert agree > 0.999 and 0.05 < valid < 0.95
    ls, lr = float(losses["self_ce"]), float(ref_losses["self_ce"])
    print("self_ce %.6f vs %.6f" % (ls, lr))
    assert abs(ls - lr) <= 2e-3 * max(abs(lr), 1e-3)
    names = [k for k, _ in m.backbone.named_parameters()]
    params = dict(m.backbone.named_parameters())
    mine = np.array([params[n].grad.double().norm().item() for n in names])
    ref = np.array([student[n].grad.double().norm().item() for n in names])
    relerr = np.abs(mine - ref) / np.maximum(ref, 1e-12)
    print("gradient norms of %d tensors: max rel err %.2e (%s)" % (len(names), relerr.max(), names[int(relerr.argmax())]))
    assert relerr.max() < 2e-2
    for n in ("model.layer5.conv2d_list.3.weight", "model.layer4.2.conv3.weight", "model.layer3.5.conv2.weight", "model.conv1.weight"):
        e = rel(params[n].grad.cpu(), student[n].grad)[0]
        print("   grad", n, "rel-L2 %.2e" % e)
        assert e < 3e-2, n           # same end-to-end bar as tests/test_step_gpu.py (label flips / ReLU flips, reasons there)


CASES = [("resnet101", 8, 3, (512, 512)),      # configs[1] / [2]: the bench line, 24 crops per GPU
         ("fcn", 4, 4, (640, 640)),            # configs[3]: 16 x 640^2 K=4 on 4 GPUs -> 4 groups per GPU
         ("resnet101", 1, 6, (1024, 1024))]    # configs[4]: 8 x 1024^2 K=6 on 8 GPUs -> 1 group per GPU


@pytest.mark.parametrize("arch,G,K,HW", CASES, ids=["cfg1_resnet101_24x512", "cfg3_fcn_16x640", "cfg4_resnet101_6x1024"])
def test_backbone_and_tail_at_bench_batch(arch, G, K, HW):
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m, cfg, sd = _net(arch)
        x, y, x2, A, Ai = [t.cuda() for t in synth.make_target_batch(G, K, HW, seed=0)]
        BT = G * K
        with torch.no_grad():
            mine = m.backbone.logits(x2)                       # the whole per-GPU batch through the tcgen05 kernels
            torch.cuda.synchronize()
            sd_gpu = {k: v.cuda() for k, v in sd.items()}
            fwd = {"resnet101": O.resnet101_logits, "fcn": O.vgg16_fcn8s_logits}[arch]
            ref = torch.cat([fwd(sd_gpu, x2[i:i + K]) for i in range(0, BT, K)])      # strict-fp32 ATen / cuDNN, K crops at a time
            l2, mx = rel(mine, ref)
            print("%s %d x %dx%d: backbone logits rel-L2 %.2e max %.2e (absmax %.1f)" % (arch, BT, HW[0], HW[1], l2, mx, ref.abs().max().item()))
            assert l2 < 1e-3 and mx < 1e-3
            # ---- tail of the whole batch from OUR teacher logits (identical logits in -> identical masks out)
            m.running_conf.fill_(0.02)
            rc0 = m.running_conf.clone()
            ws = m._tail(mine, y, A, Ai, K)
            torch.cuda.synchronize()
            ign = (y == -1)
            labels_ref = torch.empty(BT, HW[0], HW[1], dtype=torch.int64, device="cuda")
            amb = torch.zeros(BT, HW[0], HW[1], dtype=torch.bool, device="cuda")
            conf_err = 0.0
            # the running-conf update uses the mean over the WHOLE batch (sac.py:104-117): accumulate it group by group
            pbar = torch.zeros(19, device="cuda", dtype=torch.double)
            for g0 in range(0, BT, K):
                up = torch.nn.functional.interpolate(mine[g0:g0 + K], HW, mode="bilinear", align_corners=True)
                pbar += torch.softmax(up, 1).double().sum(dim=(0, 2, 3))
            pbar = (pbar / (BT * HW[0] * HW[1])).float()
            rc = rc0.clone()
            new = (pbar > 1e-8) & (rc == cfg.THRESHOLD_BETA)
            rc[new] = pbar[new]
            rc = cfg.STAT_MOMENTUM * rc + (1 - cfg.STAT_MOMENTUM) * pbar
            assert rel(m.running_conf, rc)[1] < 1e-5
            for g0 in range(0, BT, K):
                s = slice(g0, g0 + K)
                refined, _, _ = O.refine(mine[s], HW, K, A[s], Ai[s], ign[s], rc, cfg, training=False)
                lab, conf, idx, thr = O.pseudo_labels_probs(refined, ign[s], rc, cfg, cfg.CONF_DISCOUNT)
                labels_ref[s] = lab
                conf_err = max(conf_err, (ws["conf"][s] - conf).abs().max().item())
                top2 = refined.topk(2, dim=1).values
                thr_px = thr.gather(1, idx.view(K, -1)).view(K, *HW)
                amb[s] = ((conf.squeeze(1) - thr_px).abs() < 1e-5) | (((top2[:, 0] - top2[:, 1]) < 1e-5) & (conf.squeeze(1) > 0))
            mism = ws["labels"].long() != labels_ref
            valid = (labels_ref != 255).float().mean().item()
            print("tail: conf max err %.2e, label mismatches %d of %d (ambiguous pixels %d), valid fraction %.3f"
                  % (conf_err, int(mism.sum()), mism.numel(), int(amb.sum()), valid))
            assert conf_err < 5e-5
            assert int((mism & ~amb).sum()) == 0
            assert int(amb.sum()) < 1e-4 * amb.numel() and 0.02 < valid < 0.98
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
