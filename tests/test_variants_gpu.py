"""GPU parity of the non-default SAC tail variants against golden vectors from the REAL reference methods:
CONF_POOL = minentropy_pool, CONF_POOL_ON = False, LOSS = focal_ce (and the defaults on the same inputs)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G, K, HW = 2, 3, (96, 96)


@pytest.mark.parametrize("name,pool,pool_on", [("avg", "avg_pool", True), ("minent", "minentropy_pool", True), ("off", "avg_pool", False)])
def test_tail_variant_labels_and_losses(name, pool, pool_on):
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    from da_sac_b200.models.sac import _StudentLossFn
    g = np.load(os.path.join(HERE, "golden", "sac_tail_variants.npz"))
    _, y, _, A, Ai = [t.cuda() for t in synth.make_target_batch(G, K, HW, seed=21)]
    tl = torch.from_numpy(g["teacher_logits"]).cuda()
    sl = torch.from_numpy(g["student_logits"]).cuda()
    for loss in ("focal_ce_conf", "focal_ce"):
        cfg = type("Cfg", (synth.ModelCfg,), {"CONF_POOL": pool, "CONF_POOL_ON": pool_on, "LOSS": loss})()
        m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
        m.cuda().train()
        m.running_conf.copy_(torch.from_numpy(g["rc0"]).cuda())
        refined = torch.empty(G * K, 19, *HW, device="cuda")
        ws = m._tail(tl, y, A, Ai, K, refined=refined)
        torch.cuda.synchronize()
        assert torch.allclose(m.running_conf.cpu(), torch.from_numpy(g[name + "_running_conf"]), rtol=1e-5, atol=1e-8)
        assert (refined[:, :, ::3, ::3].cpu() - torch.from_numpy(g[name + "_refined_sub"])).abs().max() < 2e-5
        conf, gconf = ws["conf"].cpu(), torch.from_numpy(g[name + "_conf"])
        assert (conf - gconf).abs().max() < 2e-5
        lab, glab = ws["labels"].cpu(), torch.from_numpy(g[name + "_labels"])
        amb = torch.from_numpy(g[name + "_ambiguous"])
        mism = lab != glab
        print(name, loss, "label mismatches", int(mism.sum()), "ambiguous", int(amb.sum()))
        assert int((mism & ~amb).sum()) == 0
        # the loss on the GOLDEN labels / confidences (so a flipped ambiguous pixel cannot leak into the comparison)
        ws["labels"].copy_(glab.cuda())
        ws["conf_mean"].copy_(gconf.cuda().mean(0).view(*HW))
        _, self_ce = _StudentLossFn.apply(m, sl, y.clone(), ws)
        ref = float(g[name + "_" + loss][0])
        print("   self_ce %.7f reference %.7f" % (float(self_ce), ref))
        assert abs(float(self_ce) - ref) < 2e-5 * max(1.0, abs(ref))
