"""GPU parity of the full SAC target step (through da_sac_b200.models -> C ABI) against the CPU oracle and
the golden vectors of the real reference.  Bars (BASELINE.json north_star): logits within 1e-3 relative
(max-norm and rel-L2), pseudo-label masks bit-exact given identical teacher logits."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_GROUPS, K, HW = 2, 2, (128, 128)


def rel(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def planes_to_nchw(pl, N, H, W, Cc):
    return (pl.hi.float() + pl.lo.float()).view(N, H, W, Cc).permute(0, 3, 1, 2)


@pytest.fixture(scope="module")
def net():
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    cfg = synth.ModelCfg()
    m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    m.backbone.load_state_dict(synth.make_backbone_params(seed=123))
    m.cuda()
    m.train()
    return m, cfg


def test_backbone_forward_matches_oracle(net):
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    m, cfg = net
    x = synth.make_target_batch(N_GROUPS, K, HW, seed=0)[0]
    sd = synth.make_backbone_params(seed=123)
    taps = {}
    with torch.no_grad():
        ref = O.resnet101_logits(sd, x, taps)
    bb = m.backbone
    bb.ensure_flat(torch.device("cuda"))
    eng = bb.engine(x.shape[0], HW[0], HW[1])
    out = torch.empty(x.shape[0], 19, *eng.net["out_hw"], device="cuda")
    eng.forward(bb._flat, bb._planes(True), x.cuda(), out, keep=True)
    torch.cuda.synchronize()
    st = eng.net["stem"]
    e_stem = rel(planes_to_nchw(eng.act["stem"], x.shape[0], st.hout, st.wout, 64), taps["stem"])
    ph, pw = eng.net["pool_hw"]
    e_pool = rel(planes_to_nchw(eng.act["pool"], x.shape[0], ph, pw, 64), taps["pool"])
    errs = {"stem": e_stem, "pool": e_pool}
    for li, last in ((1, "model.layer1.2.conv3"), (2, "model.layer2.3.conv3"), (3, "model.layer3.22.conv3"), (4, "model.layer4.2.conv3")):
        s = eng.net["specs"][last]
        errs["layer%d" % li] = rel(planes_to_nchw(eng.act[last], x.shape[0], s.hout, s.wout, s.K), taps["layer%d" % li])
    errs["logits"] = rel(out, ref)
    print(errs)
    for k, (l2, mx) in errs.items():
        assert l2 < 1e-3 and mx < 1e-3, (k, l2, mx, errs)
    assert errs["logits"][1] < 2e-4      # bf16x3 should sit far inside the 1e-3 bar


def test_tail_labels_bit_exact_on_golden_teacher_logits(net, golden):
    """identical teacher logits in -> identical pseudo-label masks out (outside the audited-ambiguous pixels)"""
    from da_sac_b200 import synth
    m, cfg = net
    _, y, _, A, Ai = synth.make_target_batch(N_GROUPS, K, HW, seed=0)
    for step in (0, 1):
        pre = "s%d_" % step
        tl = torch.from_numpy(golden[pre + "teacher_logits"]).cuda()
        # running_conf before this step's update: beta at step 0, golden s0 value at step 1
        rc0 = torch.full((19,), cfg.THRESHOLD_BETA) if step == 0 else torch.from_numpy(golden["s0_running_conf"])
        m.running_conf.copy_(rc0.cuda())
        m.train()
        ws = m._tail(tl, y.cuda(), A.cuda(), Ai.cuda(), K)
        torch.cuda.synchronize()
        assert rel(m.running_conf, golden[pre + "running_conf"])[1] < 1e-5
        lab = ws["labels"].cpu()
        glab = torch.from_numpy(golden[pre + "teacher_labels"])
        conf = ws["conf"].cpu()
        gconf = torch.from_numpy(golden[pre + "teacher_conf"])
        assert (conf - gconf).abs().max() < 2e-5, (conf - gconf).abs().max()
        amb = torch.from_numpy(golden[pre + "ambiguous"])
        mism = (lab != glab)
        print("step", step, "label mismatches", int(mism.sum()), "of which ambiguous", int((mism & amb).sum()))
        assert int((mism & ~amb).sum()) == 0
        assert int(mism.sum()) <= int(amb.sum())
        # size-independent self-consistency: labels are an integer function of (conf, idx, thresholds, y)
        thr = ws["thresholds"].cpu()
        idx = ws["idx"].cpu().long().squeeze(1)
        exp = torch.where(conf.squeeze(1) > thr.gather(1, idx.view(idx.shape[0], -1)).view_as(idx), idx, torch.full_like(idx, 255))
        exp[y == -1] = 255
        assert torch.equal(exp.to(torch.uint8), lab)


def test_two_training_steps_match_reference_golden(net, golden):
    from da_sac_b200 import synth
    m, cfg = net
    m.backbone.load_state_dict(synth.make_backbone_params(seed=123))
    m.slow_init[0] = False
    m.running_conf.zero_()
    m.train()
    groups = m.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY)
    optim = torch.optim.SGD(groups, momentum=cfg.MOMENTUM)
    batch = synth.make_target_batch(N_GROUPS, K, HW, seed=0)
    names = [str(n) for n in golden["grad_names"]]
    params = dict(m.backbone.named_parameters())
    for step in (0, 1):
        x, y, x2, A, Ai = [t.clone().cuda() for t in batch]
        losses, outs = m(x, y, x2, A, Ai, use_teacher=True, update_teacher=(step == 0), T=K)
        optim.zero_grad()
        (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
        torch.cuda.synchronize()
        pre = "s%d_" % step
        assert (y.cpu() == torch.from_numpy(golden[pre + "mask_gt"]).long()).all()       # in-place -1 -> 255
        l2, mx = rel(outs["logits"].detach(), golden[pre + "logits"])
        print("step", step, "student logits rel-L2 %.2e max %.2e" % (l2, mx))
        assert l2 < 1e-3 and mx < 1e-3
        assert rel(outs["running_conf"], golden[pre + "running_conf"])[1] < 1e-4
        lab = outs["teacher_labels"].cpu().to(torch.uint8)
        glab = torch.from_numpy(golden[pre + "teacher_labels"])
        agree = (lab == glab).float().mean().item()
        print("step", step, "pseudo-label agreement end-to-end %.6f" % agree)
        assert agree > 0.999
        assert rel(outs["teacher_conf"], golden[pre + "teacher_conf"])[1] < 1e-3
        assert rel(outs["teacher_refined"][:, :, ::4, ::4], golden[pre + "teacher_refined_sub"])[1] < 1e-3
        for k in ("self_ce", "loss_ce", "teacher_diff"):
            g = float(golden[pre + k].reshape(-1)[0]); v = float(losses[k].detach().reshape(-1)[0])
            print(k, v, g)
            assert abs(v - g) <= 2e-3 * max(abs(g), 1e-3), (k, v, g)
        gn = golden[pre + "grad_norms"]
        mine = np.array([params[n].grad.double().norm().item() for n in names])
        relerr = np.abs(mine - gn) / np.maximum(gn, 1e-12)
        print("step", step, "grad-norm max rel err %.2e (%s)" % (relerr.max(), names[int(relerr.argmax())]))
        assert relerr.max() < 1e-2
        for key in golden.files:
            if key.startswith(pre + "grad::"):
                n = key.split("::")[1]
                g = params[n].grad
                g = g.flatten()[:60000] if g.numel() > 60000 else g
                e = rel(g.reshape(golden[key].shape), golden[key])[0]
                print("   grad", n, "rel-L2 %.2e" % e)
                # end-to-end the only sizeable contribution is the handful of flipped pseudo-label pixels
                # (2 of 65536 here => O(1e-2) on the deepest layers); the backward arithmetic itself is
                # checked layer by layer in test_backward_matches_reference_given_golden_pseudo_labels
                assert e < 3e-2, key
        if step == 0:
            optim.step()
            v = m.backbone.model.layer3[5].conv2.weight.detach().flatten()[:60000]
            assert rel(v, golden["s0_post_step::model.layer3.5.conv2.weight"])[1] < 1e-5


# Expected gradient agreement by depth.  Even with identical pseudo labels the comparison has a floor that no
# implementation can beat: activations agree with the reference to ~3e-5, so a fraction ~1e-5 of the ReLU units per
# layer sits on the other side of zero; each flipped unit changes the gradient through it completely, i.e. a relative
# L2 error of ~sqrt(1e-5) = 3e-3 per ReLU layer, accumulating in quadrature towards the input (33 ReLU layers ->
# ~1e-2 at conv1).  Close to the loss the agreement is 1e-4.  (The kernels themselves are checked to 2e-5 against
# fp64 in tests/test_conv_gpu.py; gradient *norms* agree to 2.4e-3 everywhere.)
GRAD_TOL = {"model.layer5": 1e-3, "model.layer4": 2e-3, "model.layer3": 8e-3, "model.layer2": 1.5e-2,
            "model.layer1": 2.5e-2, "model.conv1": 2.5e-2, "model.bn1": 2.5e-2}


def grad_tol(name):
    for k, v in GRAD_TOL.items():
        if name.startswith(k):
            return v
    return 2.5e-2


def test_backward_matches_reference_given_golden_pseudo_labels(net, golden):
    """Same two steps, but the pseudo labels / confidence / running_conf that feed the loss are overwritten with the
    reference's golden values, so that every parameter gradient can be compared without label-flip noise."""
    from da_sac_b200 import synth
    m, cfg = net
    m.backbone.load_state_dict(synth.make_backbone_params(seed=123))
    m.slow_init[0] = False
    m.running_conf.zero_()
    m.train()
    optim = torch.optim.SGD(m.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    batch = synth.make_target_batch(N_GROUPS, K, HW, seed=0)
    names = [str(n) for n in golden["grad_names"]]
    params = dict(m.backbone.named_parameters())
    orig_tail = m._tail
    state = {"step": 0}

    def tail_with_golden(*a, **k):
        ws = orig_tail(*a, **k)
        pre = "s%d_" % state["step"]
        ws["labels"].copy_(torch.from_numpy(golden[pre + "teacher_labels"]).cuda())
        conf = torch.from_numpy(golden[pre + "teacher_conf"]).cuda()
        ws["conf"].copy_(conf)
        ws["conf_mean"].copy_(conf.mean(0)[0])
        m.running_conf.copy_(torch.from_numpy(golden[pre + "running_conf"]).cuda())
        return ws

    m._tail = tail_with_golden
    try:
        for step in (0, 1):
            state["step"] = step
            x, y, x2, A, Ai = [t.clone().cuda() for t in batch]
            losses, outs = m(x, y, x2, A, Ai, use_teacher=True, update_teacher=(step == 0), T=K)
            optim.zero_grad()
            (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
            torch.cuda.synchronize()
            pre = "s%d_" % step
            g = float(golden[pre + "self_ce"].reshape(-1)[0]); v = float(losses["self_ce"].detach().reshape(-1)[0])
            assert abs(v - g) <= 5e-4 * abs(g), (v, g)
            gn = golden[pre + "grad_norms"]
            mine = np.array([params[n].grad.double().norm().item() for n in names])
            relerr = np.abs(mine - gn) / np.maximum(gn, 1e-12)
            print("step", step, "grad-norm max rel err %.2e (%s)" % (relerr.max(), names[int(relerr.argmax())]))
            assert relerr.max() < 5e-3
            for key in golden.files:
                if key.startswith(pre + "grad::"):
                    n = key.split("::")[1]
                    gg = params[n].grad
                    gg = gg.flatten()[:60000] if gg.numel() > 60000 else gg
                    e = rel(gg.reshape(golden[key].shape), golden[key])[0]
                    print("   grad", n, "rel-L2 %.2e (tol %.1e)" % (e, grad_tol(n)))
                    assert e < grad_tol(n), key
            if step == 0:
                optim.step()
    finally:
        m._tail = orig_tail


def test_vgg16_config1_two_steps_match_reference_golden():
    """BASELINE.json configs[0] (VGG-16 DeepLabv2, 1 crop 256x256, K=1) on the B200 path vs the reference's golden"""
    import os
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sac_vgg16_cfg1.npz"))
    cfg = synth.ModelCfgVGG16()
    m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    m.backbone.load_state_dict(synth.make_vgg16_params(seed=321))
    m.cuda().train()
    optim = torch.optim.SGD(m.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    batch = synth.make_target_batch(1, 1, (256, 256), seed=0)
    names = [str(n) for n in g["grad_names"]]
    params = dict(m.backbone.named_parameters())
    for step in (0, 1):
        x, y, x2, A, Ai = [t.clone().cuda() for t in batch]
        losses, outs = m(x, y, x2, A, Ai, use_teacher=True, update_teacher=(step == 0), T=1)
        optim.zero_grad()
        (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
        torch.cuda.synchronize()
        pre = "s%d_" % step
        l2, mx = rel(outs["logits"].detach(), g[pre + "logits"])
        print("vgg step", step, "logits rel-L2 %.2e max %.2e" % (l2, mx))
        assert l2 < 1e-3 and mx < 1e-3
        lab = outs["teacher_labels"].cpu().to(torch.uint8)
        agree = (lab == torch.from_numpy(g[pre + "teacher_labels"])).float().mean().item()
        print("vgg step", step, "pseudo-label agreement %.6f" % agree)
        assert agree > 0.999
        for k in ("self_ce", "loss_ce", "teacher_diff"):
            gv = float(g[pre + k].reshape(-1)[0]); v = float(losses[k].detach().reshape(-1)[0])
            print(k, v, gv)
            assert abs(v - gv) <= 3e-3 * max(abs(gv), 1e-3), (k, v, gv)
        gn = g[pre + "grad_norms"]
        mine = np.array([params[n].grad.double().norm().item() for n in names])
        relerr = np.abs(mine - gn) / np.maximum(gn, 1e-12)
        print("vgg step", step, "grad-norm max rel err %.2e (%s)" % (relerr.max(), names[int(relerr.argmax())]))
        assert relerr.max() < 2e-2
        for key in g.files:
            if key.startswith(pre + "grad::"):
                n = key.split("::")[1]
                gg = params[n].grad
                gg = gg.flatten()[:60000] if gg.numel() > 60000 else gg
                e = rel(gg.reshape(g[key].shape), g[key])[0]
                print("   grad", n, "rel-L2 %.2e" % e)
                assert e < 3e-2, key
        if step == 0:
            optim.step()
            assert rel(m.backbone.features[44].bias.detach(), g["s0_post_step::features.44.bias"])[1] < 1e-4


def test_source_pass_loss_ce_backward_matches_oracle(net):
    """Trainer.step(train=True) semantics (train.py:119-138): net(image, gt) -> loss_ce.backward(). The plain CE against y
    goes through the same fused loss kernels (labels == NULL mode)."""
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    m, cfg = net
    sd = synth.make_backbone_params(seed=123)
    m.backbone.load_state_dict(sd)
    m.train()
    x, y, _, _, _ = synth.make_target_batch(N_GROUPS, K, HW, seed=3)
    y = y.clone(); y[y == -1] = 255
    for p in m.backbone.parameters():
        p.grad = None
    losses, outs = m(x.cuda(), y.clone().cuda())
    assert "self_ce" not in losses
    losses["loss_ce"].mean().backward()
    torch.cuda.synchronize()
    student = O.as_leaf_params(sd)
    ref_losses, _ = O.backbone_forward(student, x, y)
    ref_losses["loss_ce"].mean().backward()
    assert abs(float(losses["loss_ce"].detach().reshape(-1)[0]) - float(ref_losses["loss_ce"].detach().reshape(-1)[0])) < 2e-4 * float(ref_losses["loss_ce"].detach().reshape(-1)[0])
    params = dict(m.backbone.named_parameters())
    for n in ("model.layer5.conv2d_list.0.weight", "model.layer4.1.conv2.weight", "model.layer3.10.bn2.weight", "model.layer1.0.conv1.weight"):
        e = rel(params[n].grad, student[n].grad)[0]
        print("   source-pass grad", n, "rel-L2 %.2e" % e)
        assert e < grad_tol(n), n


def test_fcn8s_two_steps_match_reference_golden():
    """VGG-16 FCN-8s (BASELINE.json configs[3] architecture; Dropout2d off for RNG-independent parity) vs the reference"""
    import os
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sac_fcn8s_tiny.npz"))
    cfg = synth.ModelCfgFCN()
    m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"), drop_rate=0.0)
    m.backbone.load_state_dict(synth.make_fcn_params(seed=213))
    m.cuda().train()
    optim = torch.optim.SGD(m.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    batch = synth.make_target_batch(1, 2, (128, 128), seed=0)
    names = [str(n) for n in g["grad_names"]]
    params = dict(m.backbone.named_parameters())
    for step in (0, 1):
        x, y, x2, A, Ai = [t.clone().cuda() for t in batch]
        losses, outs = m(x, y, x2, A, Ai, use_teacher=True, update_teacher=(step == 0), T=2)
        optim.zero_grad()
        (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
        torch.cuda.synchronize()
        pre = "s%d_" % step
        l2, mx = rel(outs["logits"].detach(), g[pre + "logits"])
        print("fcn step", step, "logits rel-L2 %.2e max %.2e" % (l2, mx))
        assert l2 < 1e-3 and mx < 1e-3
        lab = outs["teacher_labels"].cpu().to(torch.uint8)
        agree = (lab == torch.from_numpy(g[pre + "teacher_labels"])).float().mean().item()
        print("fcn step", step, "pseudo-label agreement %.6f" % agree)
        assert agree > 0.999
        for k in ("self_ce", "loss_ce", "teacher_diff"):
            gv = float(g[pre + k].reshape(-1)[0]); v = float(losses[k].detach().reshape(-1)[0])
            print(k, v, gv)
            assert abs(v - gv) <= 5e-3 * max(abs(gv), 1e-3), (k, v, gv)
        gn = g[pre + "grad_norms"]
        mine = np.array([params[n].grad.double().norm().item() for n in names])
        relerr = np.abs(mine - gn) / np.maximum(gn, 1e-12)
        print("fcn step", step, "grad-norm max rel err %.2e (%s)" % (relerr.max(), names[int(relerr.argmax())]))
        assert relerr.max() < 3e-2
        for key in g.files:
            if key.startswith(pre + "grad::"):
                n = key.split("::")[1]
                gg = params[n].grad
                gg = gg.flatten()[:60000] if gg.numel() > 60000 else gg
                e = rel(gg.reshape(g[key].shape), g[key])[0]
                print("   grad", n, "rel-L2 %.2e" % e)
                assert e < 3e-2, key
        if step == 0:
            optim.step()
            assert rel(m.backbone.vgg_head[8].bias.detach(), g["s0_post_step::vgg_head.8.bias"])[1] < 1e-4


def test_fcn8s_dropout_train_mode_runs():
    """Dropout2d active (default drop_rate): the step runs, losses are finite, every parameter receives a gradient"""
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    cfg = synth.ModelCfgFCN()
    m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    m.backbone.load_state_dict(synth.make_fcn_params(seed=213))
    m.cuda().train()
    x, y, x2, A, Ai = [t.cuda() for t in synth.make_target_batch(1, 2, (128, 128), seed=0)]
    losses, outs = m(x, y, x2, A, Ai, use_teacher=True, update_teacher=True, T=2)
    (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(v).all() for v in losses.values())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.backbone.parameters())


def test_joint_source_target_step_matches_oracle(net):
    """One ``train_epoch`` iteration with TARGET_ONLY=False (train.py:266-298): source ``loss_ce.backward()`` and target
    ``LR_TARGET*self_ce`` backward accumulate into ONE optimiser step.  Checks (a) plain autograd accumulation through the
    drop-in (two backward passes without zero_grad, what an unmodified train.py does) and (b) trainer.JointStepper
    (flat two-buffer sum, one all-reduce) against the oracle's accumulated gradients and post-SGD weights."""
    from da_sac_b200 import synth
    from da_sac_b200.trainer import JointStepper
    from oracle import sac_oracle as O
    m, cfg = net
    sd = synth.make_backbone_params(seed=123)
    G, Kk, hw = 1, 2, (96, 96)
    tgt = synth.make_target_batch(G, Kk, hw, seed=11)
    xs, ys, _, _, _ = synth.make_target_batch(1, 2, (96, 128), seed=12)      # source crops: another shape on purpose
    ys = torch.randint(0, 19, ys.shape); ys[:, :4] = 255
    # ---- oracle
    student = O.as_leaf_params(sd)
    teacher = {k: v.detach().clone() for k, v in student.items()}
    optim = torch.optim.SGD(O.parameter_groups(student, cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    optim.zero_grad()
    ls, _ = O.backbone_forward(student, xs, ys)
    ls["loss_ce"].mean().backward()
    rc = torch.full((19,), cfg.THRESHOLD_BETA)
    lt, _, _ = O.sac_target_forward(student, teacher, rc, tgt, Kk, cfg, True)
    (cfg.LR_TARGET * lt["self_ce"].mean()).backward()
    ref_grads = {k: v.grad.clone() for k, v in student.items() if v.requires_grad}
    optim.step()
    names = ("model.layer5.conv2d_list.0.weight", "model.layer4.1.conv2.weight", "model.layer3.10.bn2.weight",
             "model.layer3.10.bn2.bias", "model.layer1.0.conv1.weight", "model.conv1.weight")
    # ---- (a) autograd accumulation, as train.py drives it
    m.backbone.load_state_dict(sd); m.train(); m.slow_init[0] = False
    for p in m.backbone.parameters():
        p.grad = None
    losses_s, _ = m(xs.cuda(), ys.clone().cuda())
    losses_s["loss_ce"].mean().backward()
    x, y, x2, A, Ai = [t.clone().cuda() for t in tgt]
    losses_t, _ = m(x, y, x2, A, Ai, use_teacher=True, update_teacher=True, T=Kk)
    (cfg.LR_TARGET * losses_t["self_ce"].mean()).backward()
    torch.cuda.synchronize()
    assert abs(float(losses_s["loss_ce"]) - float(ls["loss_ce"])) < 2e-4 * abs(float(ls["loss_ce"]))
    assert abs(float(losses_t["self_ce"]) - float(lt["self_ce"])) < 5e-3 * max(abs(float(lt["self_ce"])), 1e-3)
    params = dict(m.backbone.named_parameters())
    for n in names:
        e = rel(params[n].grad, ref_grads[n])[0]
        print("   joint (autograd) grad", n, "rel-L2 %.2e" % e)
        assert e < grad_tol(n), n
    # ---- (b) JointStepper
    m.backbone.load_state_dict(sd); m.slow_init[0] = False
    for p in m.backbone.parameters():
        p.grad = None
    st = JointStepper(m, cfg, Kk, torch.device("cuda"))
    out = st.step_joint((xs.cuda(), ys.clone().cuda()), tuple(t.clone().cuda() for t in tgt), update_teacher=True)
    torch.cuda.synchronize()
    assert abs(float(out["loss_ce_source"]) - float(ls["loss_ce"])) < 2e-4 * abs(float(ls["loss_ce"]))
    gflat = m.backbone._grad
    for n in names:
        e = rel(gflat.view(n), ref_grads[n])[0]
        print("   joint (stepper) grad", n, "rel-L2 %.2e" % e)
        assert e < grad_tol(n), n
    for n in ("model.layer5.conv2d_list.1.bias", "model.layer3.5.conv2.weight", "model.bn1.weight"):
        new, ref_new, old = params[n].detach().cpu().double(), student[n].detach().double(), sd[n].double()
        # compare the UPDATE (the weights themselves agree trivially): w_new - w_old
        e = ((new - old) - (ref_new - old)).norm() / (ref_new - old).norm().clamp_min(1e-30)
        print("   joint post-SGD update", n, "rel-L2 %.2e" % float(e))
        assert float(e) < 2e-2, n


def test_training_reduces_the_loss_and_teacher_follows():
    """Does it actually train?  30 target steps on one fixed batch (device-augmented views, fused SGD): the supervised
    monitor loss against fixed pseudo ground truth is not what is optimised, so the check is on self_ce itself, on the
    teacher distance (0 right after the teacher is initialised from the student, growing while the student moves) and on
    finiteness."""
    import random
    from da_sac_b200 import augment as AUG, synth
    from da_sac_b200.models import get_model
    from da_sac_b200.trainer import TargetStepper
    cfg = type("Cfg", (synth.ModelCfg,), {"NET_MOMENTUM_ITER": 10, "LR": 2.5e-3})()
    m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    m.backbone.load_state_dict(synth.make_backbone_params(seed=123))
    m.cuda().train()
    G, Kk, hw = 2, 2, (128, 128)
    st = TargetStepper(m, cfg, Kk, torch.device("cuda"))
    random.seed(4); torch.manual_seed(4)
    g = torch.Generator().manual_seed(9)
    low = torch.rand(G, 3, 6, 6, generator=g)
    base = (torch.nn.functional.interpolate(low, hw, mode="bicubic", align_corners=False).clamp(0, 1) * 255).round()
    base = base.to(torch.uint8).permute(0, 2, 3, 1).contiguous().cuda()
    aug = AUG.TargetAugmenter(Kk, hw)
    batch = aug(base)
    hist = []
    for it in range(30):
        out = st.step(tuple(t.clone() for t in batch), read_losses=True)
        hist.append(out)
        assert all(np.isfinite(v) for v in out.values()), (it, out)
    first, last = np.mean([h["self_ce"] for h in hist[1:6]]), np.mean([h["self_ce"] for h in hist[-5:]])
    print("self_ce first/last: %.5f -> %.5f; teacher_diff at it 9/10/11: %.4f %.4f %.4f"
          % (first, last, hist[9]["teacher_diff"], hist[10]["teacher_diff"], hist[11]["teacher_diff"]))
    assert last < first, (first, last)
    assert hist[0]["teacher_diff"] == 0.0 and hist[9]["teacher_diff"] > hist[1]["teacher_diff"] > 0.0
    # replicas of the parameters stay finite and actually moved
    w0 = synth.make_backbone_params(seed=123)["model.layer5.conv2d_list.0.weight"]
    w1 = m.backbone.model.layer5.conv2d_list[0].weight.detach().cpu()
    assert torch.isfinite(w1).all() and (w1 - w0).abs().max() > 0
