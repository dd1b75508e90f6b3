"""Non-default SAC tail variants (MODEL.CONF_POOL = minentropy_pool, CONF_POOL_ON = False, LOSS = focal_ce): the oracle's
restatement against golden vectors from the REAL reference methods (tests/golden/make_golden_variants.py)."""
import os

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
G, K, HW = 2, 3, (96, 96)


def test_oracle_tail_variants_match_reference():
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    g = np.load(os.path.join(HERE, "golden", "sac_tail_variants.npz"))
    cfg = synth.ModelCfg()
    _, y, _, A, Ai = synth.make_target_batch(G, K, HW, seed=21)
    ign = y == -1
    tl, sl = torch.from_numpy(g["teacher_logits"]), torch.from_numpy(g["student_logits"])
    up = F.interpolate(sl, HW, mode="bilinear", align_corners=True)
    for name, pool in (("avg", "avg_pool"), ("minent", "minentropy_pool"), ("off", None)):
        refined, rc, _ = O.refine(tl, HW, K, A, Ai, ign, torch.from_numpy(g["rc0"]), cfg, training=True, pool=pool)
        assert torch.allclose(rc, torch.from_numpy(g[name + "_running_conf"]), rtol=1e-5, atol=1e-8)
        assert (refined[:, :, ::3, ::3] - torch.from_numpy(g[name + "_refined_sub"])).abs().max() < 2e-6
        labels, conf, _, _ = O.pseudo_labels_probs(refined, ign, rc, cfg, cfg.CONF_DISCOUNT)
        gold = torch.from_numpy(g[name + "_labels"].astype(np.int64))
        amb = torch.from_numpy(g[name + "_ambiguous"])
        assert int(((labels != gold) & ~amb).sum()) == 0, name
        l_conf = O.focal_ce_conf(up, gold, torch.from_numpy(g[name + "_conf"]), rc, cfg.FOCAL_P)
        l_plain = O.focal_ce(up, gold, rc, cfg.FOCAL_P)
        assert abs(float(l_conf) - float(g[name + "_focal_ce_conf"][0])) < 1e-6
        assert abs(float(l_plain) - float(g[name + "_focal_ce"][0])) < 1e-6
