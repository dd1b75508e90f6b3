"""The no-grad / eval side of the model contract: what ``train.py``'s validation loop and ``infer_val.py`` call.

  * ``net(x)`` -> ``(logits, logits_up)`` from the student, ``net(x, teacher=True)`` from the momentum network
    (/root/reference/models/sac.py:324-329);
  * the validation step (train.py:371-399): ``net.eval()``, ``torch.no_grad()``, the full ``forward`` with labels and
    ``use_teacher=True``; it reads ``logits_up``, ``teacher_init``, ``teacher_refined``, ``teacher_labels`` from ``net_outs`` and
    must leave ``running_conf`` and the teacher alone (``_update_running_conf`` only runs in training mode, sac.py:278-279).
Compared with the CPU oracle on the same seeded inputs; bars as everywhere: logits 1e-3, masks identical outside the pixels that
sit on a threshold."""
import pytest
import torch

pytestmark = pytest.mark.gpu

G, K, HW = 2, 2, (128, 128)


def rel(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def setup():
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    from oracle import sac_oracle as O
    cfg = synth.ModelCfg()
    sd = synth.make_backbone_params(seed=123)
    net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    net.backbone.load_state_dict(sd)
    net.cuda().train()
    batch = synth.make_target_batch(G, K, HW, seed=0)
    # one training step first: initialises the teacher (:= student), moves running_conf off its initial value and the student
    # off the teacher, so that "student" and "teacher" below are different networks
    x, y, x2, A, Ai = [t.clone().cuda() for t in batch]
    optim = torch.optim.SGD(net.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    losses, _ = net(x, y, x2, A, Ai, use_teacher=True, update_teacher=True, T=K)
    optim.zero_grad()
    (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
    optim.step()
    torch.cuda.synchronize()
    student = {k: v.detach().cpu().clone() for k, v in net.backbone.state_dict().items()}
    teacher = {k: v.detach().cpu().clone() for k, v in net.slow_net.state_dict().items()}
    assert rel(student["model.layer5.conv2d_list.0.weight"], teacher["model.layer5.conv2d_list.0.weight"])[0] > 1e-6
    return net, cfg, batch, student, teacher, O


def test_inference_forward_student_and_teacher(setup):
    net, cfg, batch, student, teacher, O = setup
    x = batch[0]
    net.eval()
    with torch.no_grad():
        for which, params in ((False, student), (True, teacher)):
            logits, up = net(x.cuda(), teacher=which)                    # sac.py:324-329
            ref_logits, ref_up = O.backbone_forward(params, x)
            assert tuple(up.shape) == (x.shape[0], 19) + HW and up.is_contiguous()
            e1, e2 = rel(logits, ref_logits), rel(up, ref_up)
            print("teacher=%s: logits rel-L2 %.2e max %.2e, logits_up max %.2e" % (which, e1[0], e1[1], e2[1]))
            assert max(e1) < 1e-3 and max(e2) < 1e-3
    net.train()


def test_validation_step_in_eval_mode(setup):
    net, cfg, batch, student, teacher, O = setup
    net.eval()
    rc0 = net.running_conf.detach().clone()
    t0 = net.slow_net._flat.buf.clone()
    x, y, x2, A, Ai = [t.clone().cuda() for t in batch]
    with torch.no_grad():
        losses, outs = net(x, y, x2, A, Ai, use_teacher=True, update_teacher=False, T=K)        # train.py:380-386
        got = {k: outs[k] for k in ("logits_up", "teacher_init", "teacher_refined", "teacher_labels", "teacher_conf")}
    torch.cuda.synchronize()
    assert torch.equal(net.running_conf, rc0), "eval mode must not update running_conf (sac.py:278)"
    assert torch.equal(net.slow_net._flat.buf, t0), "update_teacher=False must not move the teacher"
    assert (y != -1).all()                                                   # in place: -1 -> 255 (sac.py:337-338)
    ref_losses, ref, _ = O.sac_target_forward(student, teacher, rc0.cpu(), batch, K, cfg, training=False)
    for k in ("logits_up", "teacher_init", "teacher_refined", "teacher_conf"):
        e = rel(got[k], ref[k])
        print("%-16s rel-L2 %.2e max %.2e" % (k, e[0], e[1]))
        assert max(e) < 1e-3, (k, e)
    lab, rlab = got["teacher_labels"].cpu(), ref["teacher_labels"]
    assert lab.dtype == torch.int64 and tuple(lab.shape) == (G * K,) + HW
    # end to end the teacher logits differ by ~1e-5 from the oracle's, the confidences by up to 5e-5 (measured), and a threshold
    # moves with the peak confidence of its class: pixels within 2e-4 of their threshold (or of a tie) may fall on either side
    conf, idx, thr = ref["teacher_conf"].squeeze(1), ref["teacher_idx"].squeeze(1), ref["thresholds"]
    thr_px = thr.gather(1, idx.view(idx.shape[0], -1)).view_as(conf)
    top2 = ref["teacher_refined"].topk(2, dim=1).values
    amb = ((conf - thr_px).abs() < 2e-4) | (((top2[:, 0] - top2[:, 1]) < 2e-4) & (conf > 0))
    mism = lab != rlab
    print("label mismatches %d (ambiguous pixels %d), valid fraction %.3f" % (int(mism.sum()), int(amb.sum()), (rlab != 255).float().mean().item()))
    assert int((mism & ~amb).sum()) == 0 and (~mism).float().mean().item() > 0.9999
    for k in ("self_ce", "loss_ce"):
        v, g = float(losses[k]), float(ref_losses[k])
        assert abs(v - g) <= 2e-3 * max(abs(g), 1e-3), (k, v, g)
    net.train()
