"""Golden vectors for the NON-DEFAULT variants of the SAC tail, produced by the REAL reference methods:
MODEL.CONF_POOL = minentropy_pool (models/sac.py:218-236), MODEL.CONF_POOL_ON = False (_refine(pool=False), :284-285) and
MODEL.LOSS = focal_ce (:119-132), next to the default avg_pool / focal_ce_conf for the same inputs.

    python tests/golden/make_golden_variants.py          # build container only (/root/reference needed)

The methods are called directly on synthetic teacher / student logits (no backbone forward), so the file pins exactly the
tail arithmetic.  Writes tests/golden/sac_tail_variants.npz."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
G, K, HW, hw = 2, 3, (96, 96), (13, 13)


def main():
    from da_sac_b200 import synth
    sys.path.insert(0, REF)
    from core.config import cfg, cfg_from_file, cfg_from_list
    cfg_from_file(os.path.join(REF, "configs/deeplabv2_resnet101_train.yaml"))
    cfg_from_list(["TRAIN.GROUP_SIZE", str(K), "TRAIN.NUM_GROUPS", str(G), "DATASET.CROP_SIZE", "(%d,%d)" % HW, "MODEL.INIT_MODEL", ""])
    from models import get_model
    net = get_model(cfg.MODEL, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    sys.path.remove(REF)
    net.train()
    torch.manual_seed(11)
    x, y, x2, A, Ai = synth.make_target_batch(G, K, HW, seed=21)
    tl = torch.randn(G * K, 19, *hw) * 3.0
    tl[:, :6] += 1.5                                   # a few dominant classes, like a trained teacher
    sl = torch.randn(G * K, 19, *hw) * 2.0
    rc0 = torch.rand(19) * 0.1 + 0.005
    ign = (y == -1)
    out = {"teacher_logits": tl.numpy(), "student_logits": sl.numpy(), "rc0": rc0.numpy()}
    up = F.interpolate(sl, HW, mode="bilinear", align_corners=True)
    for name, pool_func, pool_on in (("avg", net._avg_pool, True), ("minent", net._minentropy_pool, True), ("off", net._avg_pool, False)):
        net.pool_func = pool_func
        net.running_conf.copy_(rc0)
        with torch.no_grad():
            probs, _ = net._refine(x2, tl.clone(), K, A, Ai, ign, pool=pool_on, debug=False)
            labels, conf, idx = net._pseudo_labels_probs(probs, ign, cfg.MODEL.CONF_DISCOUNT)
            l_conf, _ = net._focal_ce_conf(up, labels, conf, cfg.MODEL.FOCAL_P)
            l_plain, _ = net._focal_ce(up, labels, conf, cfg.MODEL.FOCAL_P)
        top2 = probs.topk(2, dim=1).values
        B, C = probs.shape[:2]
        cf, ix = probs.max(1)
        peaks = torch.zeros_like(probs).scatter_(1, ix[:, None], cf[:, None]).view(B, C, -1).max(-1).values
        thr = (peaks * cfg.MODEL.RUN_CONF_UPPER * (1 - torch.exp(-net.running_conf / cfg.MODEL.THRESHOLD_BETA)).view(1, C)).clamp(cfg.MODEL.RUN_CONF_LOWER)
        thr_px = thr.gather(1, ix.view(B, -1)).view_as(cf)
        amb = ((cf - thr_px).abs() < 1e-5) | (((top2[:, 0] - top2[:, 1]) < 1e-5) & (cf > 0))
        out.update({name + "_labels": labels.numpy().astype(np.uint8), name + "_conf": conf.numpy(), name + "_ambiguous": amb.numpy(),
                    name + "_refined_sub": probs[:, :, ::3, ::3].contiguous().numpy(), name + "_running_conf": net.running_conf.clone().numpy(),
                    name + "_focal_ce_conf": l_conf.mean().view(1).numpy(), name + "_focal_ce": l_plain.mean().view(1).numpy()})
        print(name, "valid %.3f" % (labels != 255).float().mean().item(), "ambiguous", int(amb.sum()),
              "focal_ce_conf %.6f focal_ce %.6f" % (float(l_conf.mean()), float(l_plain.mean())))
    path = os.path.join(HERE, "sac_tail_variants.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
