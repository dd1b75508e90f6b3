"""Target-view augmentation, CPU side: (1) the host parameter draw consumes ``random`` / torch's generator exactly like the
reference's transforms (affine operators bit-equal to DataTarget._get_affine/_get_affine_inv), (2) the numpy oracle
(oracle/aug_oracle.py) against golden vectors produced by the REAL reference PIL pipeline
(tests/golden/make_golden_aug.py): masks / labels exact, pixels to about one 8-bit grey level."""
import os
import random

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    return np.load(os.path.join(HERE, "golden", "aug_reference.npz"))


def cases(g):
    for ci in range(int(g["n_cases"])):
        yield ci, {k[len("c%d_" % ci):]: g[k] for k in g.files if k.startswith("c%d_" % ci)}


def test_parameter_draw_reproduces_reference_affine_operators():
    from da_sac_b200 import augment as AUG
    g = golden()
    for ci, c in cases(g):
        K, hw = int(c["K"]), c["base"].shape[:2]
        cfg = type("Cfg", (AUG.AugCfg,), {"RND_ZOOM": tuple(float(v) for v in c["zoom"])})
        random.seed(int(c["seed"])); torch.manual_seed(int(c["seed"]))
        rows, aff = AUG.draw_group_params(K, hw, cfg)
        assert np.array_equal(np.asarray(aff, np.float64), c["affine_params"]), ci
        A, Ai = AUG.affine_from_params(aff, hw)
        assert np.array_equal(A.numpy(), c["affine"]) and np.array_equal(Ai.numpy(), c["affine_inv"]), ci
        assert np.array_equal(np.asarray(rows, np.float32), c["rows"]), ci
        # the crop window and the affine operator describe the same map: window centre offset == (dy, dx), size == crop / s
        for r, p in zip(rows, aff):
            assert abs((r[1] + r[3] / 2 - hw[0] / 2) - p[0]) < 1e-9 and abs((r[2] + r[4] / 2 - hw[1] / 2) - p[1]) < 1e-9


def test_oracle_matches_reference_pil_pipeline():
    from da_sac_b200 import augment as AUG
    from oracle import aug_oracle as AO
    g = golden()
    std = np.asarray(AUG.STD, np.float32).reshape(1, 3, 1, 1)
    for ci, c in cases(g):
        f1, gt, f2, _, _ = AO.augment_group(c["base"], c["base_mask"], c["base_label"], c["rows"], AUG.MEAN, AUG.STD)
        ref_gt = c["gt"].astype(np.int64)
        assert np.array_equal(gt, ref_gt), (ci, int((gt != ref_gt).sum()))
        # pixel differences in 8-bit grey levels
        d2 = np.abs(f2 - c["frames2"]) * std * 255.0
        d1 = np.abs(f1 - c["frames1"]) * std * 255.0
        print("case", ci, "clean: max %.2f mean %.3f levels; noisy: max %.2f mean %.3f p99 %.2f levels"
              % (d2.max(), d2.mean(), d1.max(), d1.mean(), np.percentile(d1, 99)))
        assert d2.max() <= 1.01 and d2.mean() < 0.2, ci          # geometry: Pillow rounds between its two passes
        assert d1.mean() < 1.5 and np.percentile(d1, 99) < 6.0, ci   # blur is a true Gaussian, Pillow uses 3 box passes


def test_identity_view_is_a_pure_normalisation():
    from da_sac_b200 import augment as AUG
    from oracle import aug_oracle as AO
    g = golden()
    _, c = next(cases(g))
    H, W, _ = c["base"].shape
    row = np.zeros(16, np.float32); row[0] = 1; row[3] = H; row[4] = W
    f1, gt, f2, _, raw = AO.augment_group(c["base"], c["base_mask"], None, [row], AUG.MEAN, AUG.STD)
    assert np.array_equal(raw[0], c["base"])
    assert np.array_equal(f1, f2)
    assert set(np.unique(gt)) <= {-1, 255} and np.array_equal(gt[0] == -1, c["base_mask"] > 0)
