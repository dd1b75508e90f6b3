"""Ragged and degenerate inputs of the target step -- the golden fixtures and the bench geometries are all square crops whose
sides are multiples of 32.  Everything here is compared with ``oracle/sac_oracle.py`` (pinned to the real reference,
tests/test_oracle_golden.py) on the same seeded inputs at sizes the CPU finishes in a few seconds:

  * crops whose height and width differ and are NOT multiples of the network stride (the reference's ``ceil_mode`` max-pool and
    the stride-2 convs round differently, /root/reference/models/deeplabv2.py:125-131; the 1x1 stride-2 downsample and the
    ``align_corners=True`` up-sampling of /root/reference/models/deeplabv2.py:215-217 see odd sizes);
  * a view-group of ONE view (K = 1: the pooled teacher is the view itself, /root/reference/models/sac.py:238-305);
  * a batch in which EVERY pixel carries the padding marker -1 (/root/reference/models/sac.py:337-338): all pseudo labels are
    255, the loss is the reference's value for an empty selection and no gradient may be NaN;
  * label maps that arrive as int32 / on the wrong device are refused loudly instead of being reinterpreted.
Bars as everywhere: logits 1e-3 (max-norm and rel-L2), labels identical outside pixels within 2e-4 of a threshold or tie."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def env():
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    from oracle import sac_oracle as O
    cfg = synth.ModelCfg()
    sd = synth.make_backbone_params(seed=123)
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))

    def fresh():
        net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
        net.backbone.load_state_dict(sd)
        return net.cuda().train()
    return cfg, sd, fresh, O


def _one_step(env, batch, K):
    """one training step here and in the oracle from identical weights; returns everything the checks need"""
    cfg, sd, fresh, O = env
    net = fresh()
    x, y, x2, A, Ai = [t.clone().cuda() for t in batch]
    losses, outs = net(x, y, x2, A, Ai, use_teacher=True, update_teacher=True, T=K)
    (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
    torch.cuda.synchronize()
    student = O.as_leaf_params(sd)
    teacher = {k: v.detach().clone() for k, v in student.items()}
    rc = torch.full((19,), cfg.THRESHOLD_BETA)
    ref_losses, ref_outs, _ = O.sac_target_step(student, teacher, rc, [t.clone() for t in batch], K, cfg, optim=None)
    return net, losses, outs, student, ref_losses, ref_outs


def _check_labels(outs, ref_outs, min_agree=0.999):
    lab, rlab = outs["teacher_labels"].cpu(), ref_outs["teacher_labels"]
    assert lab.dtype == torch.int64 and lab.shape == rlab.shape
    conf, idx, thr = ref_outs["teacher_conf"].squeeze(1), ref_outs["teacher_idx"].squeeze(1), ref_outs["thresholds"]
    thr_px = thr.gather(1, idx.view(idx.shape[0], -1)).view_as(conf)
    top2 = ref_outs["teacher_refined"].topk(2, dim=1).values
    amb = ((conf - thr_px).abs() < 2e-4) | (((top2[:, 0] - top2[:, 1]) < 2e-4) & (conf > 0))
    mism = lab != rlab
    agree = (~mism).float().mean().item()
    print("labels: %d mismatches (ambiguous pixels %d), agreement %.6f, valid fraction %.3f"
          % (int(mism.sum()), int(amb.sum()), agree, (rlab != 255).float().mean().item()))
    assert int((mism & ~amb).sum()) == 0 and agree >= min_agree


def _check_grads(net, student, bar_norm=2e-2, bar_sel=3e-2):
    params = dict(net.backbone.named_parameters())
    names = list(params)
    mine = np.array([params[n].grad.double().norm().item() for n in names])
    ref = np.array([student[n].grad.double().norm().item() for n in names])
    relerr = np.abs(mine - ref) / np.maximum(ref, 1e-12)
    print("gradient norms of %d tensors: max rel err %.2e (%s)" % (len(names), relerr.max(), names[int(relerr.argmax())]))
    assert relerr.max() < bar_norm
    for n in ("model.layer5.conv2d_list.3.weight", "model.layer4.2.conv3.weight", "model.layer2.0.downsample.0.weight", "model.conv1.weight"):
        e = rel(params[n].grad, student[n].grad)[0]
        print("   grad", n, "rel-L2 %.2e" % e)
        assert e < bar_sel, n            # end-to-end bar of tests/test_step_gpu.py (label / ReLU flips; reasons there)


RAGGED = [(2, (96, 160)),       # H != W, both multiples of 32: feature map 13 x 21
          (2, (97, 131)),       # odd sides: every stride-2 stage and the ceil_mode pool round (features 13 x 17)
          (3, (70, 203))]       # not a multiple of anything, three views


@pytest.mark.parametrize("K,HW", RAGGED, ids=["96x160", "97x131", "70x203_K3"])
def test_ragged_crop_full_step_vs_cpu_oracle(env, K, HW):
    from da_sac_b200 import synth
    batch = synth.make_target_batch(1, K, HW, seed=3)
    net, losses, outs, student, ref_losses, ref_outs = _one_step(env, batch, K)
    assert tuple(outs["logits"].shape) == tuple(ref_outs["logits"].shape)
    l2, mx = rel(outs["logits"].detach(), ref_outs["logits"].detach())
    print("%dx%d K=%d: feature map %s, logits rel-L2 %.2e max %.2e" % (HW[0], HW[1], K, tuple(outs["logits"].shape[2:]), l2, mx))
    assert l2 < 1e-3 and mx < 1e-3
    e = rel(outs["logits_up"].detach(), ref_outs["logits_up"].detach())
    assert tuple(outs["logits_up"].shape[2:]) == HW and max(e) < 1e-3
    _check_labels(outs, ref_outs)
    for k in ("self_ce", "loss_ce"):
        v, g = float(losses[k].detach()), float(ref_losses[k].detach())
        print("%s %.6f vs %.6f" % (k, v, g))
        assert abs(v - g) <= 2e-3 * max(abs(g), 1e-3), (k, v, g)
    assert rel(net.running_conf, ref_outs["running_conf"])[1] < 1e-4
    _check_grads(net, student)


def test_single_view_group(env):
    """K = 1: one clean view per group, the noisy view is trained on the labels of its own clean twin"""
    from da_sac_b200 import synth
    batch = synth.make_target_batch(2, 1, (128, 96), seed=5)
    net, losses, outs, student, ref_losses, ref_outs = _one_step(env, batch, 1)
    l2, mx = rel(outs["logits"].detach(), ref_outs["logits"].detach())
    assert l2 < 1e-3 and mx < 1e-3
    _check_labels(outs, ref_outs)
    v, g = float(losses["self_ce"].detach()), float(ref_losses["self_ce"].detach())
    assert abs(v - g) <= 2e-3 * max(abs(g), 1e-3)
    _check_grads(net, student)


def test_every_pixel_ignored(env):
    """all of y is the padding marker: nothing is selected; whatever the reference returns for that (the oracle tells), no NaN
    may enter the gradients and running_conf must follow the reference's masked mean"""
    from da_sac_b200 import synth
    batch = list(synth.make_target_batch(1, 2, (96, 96), seed=7))
    batch[1] = torch.full_like(batch[1], -1)
    net, losses, outs, student, ref_losses, ref_outs = _one_step(env, batch, 2)
    lab = outs["teacher_labels"].cpu()
    assert torch.equal(lab, ref_outs["teacher_labels"]) and bool((lab == 255).all())
    v, g = float(losses["self_ce"].detach()), float(ref_losses["self_ce"].detach())
    print("self_ce with nothing selected: %r (reference %r)" % (v, g))
    assert (np.isnan(v) and np.isnan(g)) or abs(v - g) <= 1e-6
    rc, rrc = net.running_conf.cpu(), ref_outs["running_conf"]
    assert torch.equal(torch.isnan(rc), torch.isnan(rrc)) and rel(torch.nan_to_num(rc), torch.nan_to_num(rrc))[1] < 1e-4
    for n, p in net.backbone.named_parameters():
        gr, rg = p.grad, student[n].grad
        assert gr is not None and bool(torch.isfinite(gr).all()) == bool(torch.isfinite(rg).all()), n
        if bool(torch.isfinite(rg).all()):
            assert float(gr.abs().max()) <= 1e-12 + 1e-3 * float(rg.abs().max()) or rel(gr, rg)[0] < 3e-2, n


def test_wrong_label_type_is_refused(env):
    from da_sac_b200 import synth
    cfg, sd, fresh, O = env
    net = fresh()
    x, y, x2, A, Ai = [t.cuda() for t in synth.make_target_batch(1, 2, (96, 96), seed=0)]
    with pytest.raises(TypeError):
        net(x, y.int(), x2, A, Ai, use_teacher=True, update_teacher=True, T=2)
