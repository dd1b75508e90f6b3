"""GPU parity of the device-side target-view augmentation (sacb_target_augment) against the numpy oracle (level by level)
and against golden vectors from the REAL reference PIL pipeline; plus size-independent properties at BASELINE size
(8 groups x K=3 x 512x512): identity view, mask/label rules, and agreement between the pixels and the affine operators
that SAC._refine later uses to warp the views back (sac.py:289-290)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    g = np.load(os.path.join(HERE, "golden", "aug_reference.npz"))
    for ci in range(int(g["n_cases"])):
        yield ci, {k[len("c%d_" % ci):]: g[k] for k in g.files if k.startswith("c%d_" % ci)}


def test_augment_matches_oracle_and_reference_golden():
    from da_sac_b200 import augment as AUG
    from oracle import aug_oracle as AO
    std = np.asarray(AUG.STD, np.float32).reshape(1, 3, 1, 1)
    for ci, c in cases():
        K, hw = int(c["K"]), c["base"].shape[:2]
        aug = AUG.TargetAugmenter(K, hw)
        rows = torch.from_numpy(c["rows"])
        A, Ai = torch.from_numpy(c["affine"]), torch.from_numpy(c["affine_inv"])
        base = torch.from_numpy(c["base"])[None].cuda()
        f1, gt, f2, A_d, Ai_d = aug(base, torch.from_numpy(c["base_mask"])[None].cuda(),
                                    torch.from_numpy(c["base_label"])[None].cuda(), params=(rows, A, Ai))
        torch.cuda.synchronize()
        o1, ogt, o2, olev, oraw = AO.augment_group(c["base"], c["base_mask"], c["base_label"], c["rows"], AUG.MEAN, AUG.STD)
        # ---- vs the oracle: integer work exact, 8-bit levels identical up to expf ulps at rounding boundaries
        assert np.array_equal(gt.cpu().numpy(), ogt), ci
        ws = aug._ws[(1, base.device)]
        assert np.array_equal(ws["raw"].cpu().numpy(), oraw), ci
        assert np.array_equal(f2.cpu().numpy(), o2), ci
        dl = np.rint(np.abs(f1.cpu().numpy() - o1) * std * 255.0)            # noisy copy, in 8-bit grey levels
        frac_same = float((dl == 0).mean())
        print("case", ci, "noisy levels identical: %.5f, max diff %.0f" % (frac_same, dl.max()))
        assert frac_same > 0.995, ci
        assert np.percentile(dl, 99.9) <= 1.0, ci
        # ---- vs the reference's PIL pipeline
        assert np.array_equal(gt.cpu().numpy(), c["gt"].astype(np.int64)), ci
        d2 = np.abs(f2.cpu().numpy() - c["frames2"]) * std * 255.0
        d1 = np.abs(f1.cpu().numpy() - c["frames1"]) * std * 255.0
        assert d2.max() <= 1.01 and d2.mean() < 0.2, ci
        assert d1.mean() < 1.5 and np.percentile(d1, 99) < 6.0, ci
        assert torch.equal(A_d.cpu(), A) and torch.equal(Ai_d.cpu(), Ai)


def _smooth_base(G, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(G, 3, H // 32 + 2, W // 32 + 2, generator=g)
    img = torch.nn.functional.interpolate(low, (H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    return (img * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def test_augment_properties_at_baseline_size():
    from da_sac_b200 import augment as AUG
    G, K, H, W = 8, 3, 512, 512
    random.seed(3); torch.manual_seed(3)
    base = _smooth_base(G, H, W, 0).cuda()
    mask = torch.zeros(G, H, W, dtype=torch.uint8, device="cuda"); mask[:, :, W - 16:] = 1
    aug = AUG.TargetAugmenter(K, (H, W))
    f1, gt, f2, A, Ai = aug(base, mask, None)
    torch.cuda.synchronize()
    rows = aug.last_params[0]
    assert f1.shape == (G * K, 3, H, W) and gt.shape == (G * K, H, W) and A.shape == (G * K, 2, 3)
    mean = torch.tensor(AUG.MEAN, device="cuda").view(1, 3, 1, 1); std = torch.tensor(AUG.STD, device="cuda").view(1, 3, 1, 1)
    norm = (base.permute(0, 3, 1, 2).float() / 255 - mean) / std * (1 - mask[:, None].float())
    for gi in range(G):
        v0 = gi * K                                                   # view 0: un-zoomed original (maybe flipped)
        ref = norm[gi].flip(-1) if rows[v0, 0] < 0 else norm[gi]
        assert torch.allclose(f2[v0], ref, rtol=0, atol=2e-6)      # torch divides by a scalar via its reciprocal: 1 ulp
    # labels: unlabelled target data -> 255 everywhere except -1 inside padding; padded pixels are zero in both copies
    assert set(torch.unique(gt).tolist()) <= {-1, 255}
    pad = (gt == -1)[:, None].expand_as(f1)
    assert float(f1[pad].abs().max()) == 0.0 and float(f2[pad].abs().max()) == 0.0
    # geometry vs affine operators: warping every clean view to the reference frame must give back the base crop
    grid = torch.nn.functional.affine_grid(A, list(f2.shape), align_corners=False)
    aligned = torch.nn.functional.grid_sample(f2, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    valid = torch.nn.functional.grid_sample(torch.ones_like(f2[:, :1]), grid, align_corners=False)
    unmasked = torch.nn.functional.grid_sample((gt != -1)[:, None].float(), grid, align_corners=False)
    ok = ((valid > 0.999) & (unmasked > 0.999)).expand_as(f2)
    ref = norm.repeat_interleave(K, 0)
    err = ((aligned - ref).abs() * ok).sum() / ok.sum()
    print("mean |warp(view, affine) - base| inside the valid region: %.4f (normalised units)" % float(err))
    assert float(err) < 0.03
    assert float(ok.float().mean()) > 0.2
    # and the inverse operator maps the reference frame onto each view
    grid_inv = torch.nn.functional.affine_grid(Ai, list(f2.shape), align_corners=False)
    back = torch.nn.functional.grid_sample(ref, grid_inv, align_corners=False)
    inside = (gt != -1)[:, None].expand_as(f2)
    err2 = ((back - f2).abs() * inside).sum() / inside.sum()
    print("mean |warp(base, affine_inv) - view|: %.4f" % float(err2))
    assert float(err2) < 0.03


def test_step_runs_from_device_augmented_views():
    """base crops -> sacb_target_augment -> SAC target step (the chain the north star describes)"""
    from da_sac_b200 import augment as AUG, synth
    from da_sac_b200.models import get_model
    from da_sac_b200.trainer import TargetStepper
    cfg = synth.ModelCfg()
    net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    net.backbone.load_state_dict(synth.make_backbone_params(seed=123))
    net.cuda().train()
    G, K, hw = 1, 3, (128, 128)
    random.seed(0); torch.manual_seed(0)
    aug = AUG.TargetAugmenter(K, hw)
    st = TargetStepper(net, cfg, K, torch.device("cuda"))
    out = st.step(aug(_smooth_base(G, hw[0], hw[1], 1).cuda()), update_teacher=True, read_losses=True)
    assert np.isfinite(out["self_ce"]) and np.isfinite(out["loss_ce"])
