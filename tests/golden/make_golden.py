"""Generate golden vectors by running the REAL reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

    python tests/golden/make_golden.py vgg16      # BASELINE.json configs[0]: VGG-16 DeepLabv2, 1 crop 256x256, K=1

Writes tests/golden/sac_resnet101_tiny.npz (and sac_vgg16_cfg1.npz).  Protocol (mirrors
/root/reference/train.py:266-298 with TRAIN.TARGET_ONLY=True):

  step 0: update_teacher=True  -> teacher := student, running_conf := beta
          (models/sac.py:75-81), forward, LR_TARGET*self_ce backward, SGD step
  step 1: update_teacher=False -> forward, backward (no optimiser step)

Weights / inputs come from da_sac_b200.synth (seeded; regenerated identically in
the tests).  Margins of every thresholding / argmax decision are audited and
pixels closer than 1e-5 to a decision boundary are recorded in ``ambiguous``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from da_sac_b200 import synth  # noqa: E402

N_GROUPS, K, HW = 2, 2, (128, 128)
ARCH = "resnet101"
if len(sys.argv) > 1 and sys.argv[1] == "vgg16":
    # BASELINE.json configs[0]: VGG-16 DeepLabv2, 1 target crop 256x256, K=1 (the reference's CPU-runnable plumbing case)
    N_GROUPS, K, HW, ARCH = 1, 1, (256, 256), "vgg16"


if len(sys.argv) > 1 and sys.argv[1] == "fcn":
    # VGG-16 FCN-8s (BASELINE.json configs[3] architecture) at a CPU-sized crop; Dropout2d disabled (drop_rate=0) so that
    # the comparison does not depend on the RNG stream
    N_GROUPS, K, HW, ARCH = 1, 2, (128, 128), "fcn"


def build_reference_net():
    sys.path.insert(0, REF)
    from core.config import cfg, cfg_from_file, cfg_from_list
    cfg_from_file(os.path.join(REF, {"resnet101": "configs/deeplabv2_resnet101_train.yaml", "vgg16": "configs/deeplabv2_vgg16_train.yaml",
                                     "fcn": "configs/fcn_vgg16_train.yaml"}[ARCH]))
    cfg_from_list(["TRAIN.GROUP_SIZE", str(K), "TRAIN.NUM_GROUPS", str(N_GROUPS),
                   "DATASET.CROP_SIZE", "(%d,%d)" % HW, "MODEL.INIT_MODEL", ""])
    from models import get_model
    extra = {"drop_rate": 0.0} if ARCH == "fcn" else {}
    net = get_model(cfg.MODEL, 0, num_classes=19,
                    criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"), **extra)
    sys.path.remove(REF)
    return net, cfg


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    net, cfg = build_reference_net()
    sd = {"resnet101": lambda: synth.make_backbone_params(seed=123), "vgg16": lambda: synth.make_vgg16_params(seed=321),
          "fcn": lambda: synth.make_fcn_params(seed=213)}[ARCH]()
    missing = net.backbone.load_state_dict(sd, strict=True)
    print("loaded", missing)
    net.train()
    groups = net.parameter_groups(cfg.MODEL.LR, cfg.MODEL.WEIGHT_DECAY)
    optim = torch.optim.SGD(groups, momentum=cfg.MODEL.MOMENTUM)

    batch = synth.make_target_batch(N_GROUPS, K, HW, seed=0)
    out = {}
    for step in (0, 1):
        x, y, x2, A, Ai = [t.clone() for t in batch]
        losses, outs = net(x, y, x2, A, Ai, use_teacher=True, update_teacher=(step == 0), T=K)
        optim.zero_grad()
        (cfg.MODEL.LR_TARGET * losses["self_ce"].mean()).backward()
        pre = "s%d_" % step
        if "logits" in outs:
            out[pre + "logits"] = outs["logits"].detach().numpy()
        else:                                   # VGG16_FCN8s exposes logits_up only (fcn.py:149)
            with torch.no_grad():
                out[pre + "logits"] = net.backbone._backbone(x).numpy()
        with torch.no_grad():
            tl, _ = net.slow_net(x2)
        out[pre + "teacher_logits"] = tl.numpy()
        refined = outs["teacher_refined"]
        out[pre + "teacher_refined_sub"] = refined[:, :, ::4, ::4].contiguous().numpy()
        out[pre + "teacher_refined_sum"] = refined.double().sum(dim=(2, 3)).numpy()
        out[pre + "teacher_conf"] = outs["teacher_conf"].numpy()
        out[pre + "teacher_labels"] = outs["teacher_labels"].numpy().astype(np.uint8)
        out[pre + "running_conf"] = outs["running_conf"].clone().numpy()
        out[pre + "mask_gt"] = y.numpy().astype(np.uint8)
        for k in ("self_ce", "loss_ce", "teacher_diff"):
            out[pre + k] = losses[k].detach().numpy()
        # margin audit (SURVEY.md section 7 "bit-exact masks")
        conf, idx = refined.max(1)
        top2 = refined.topk(2, dim=1).values
        gap = top2[:, 0] - top2[:, 1]
        B, C = refined.shape[:2]
        peaks = torch.zeros_like(refined).scatter_(1, idx[:, None], conf[:, None]).view(B, C, -1).max(-1).values
        thr = peaks * cfg.MODEL.RUN_CONF_UPPER * (1 - torch.exp(-outs["running_conf"] / cfg.MODEL.THRESHOLD_BETA)).view(1, C)
        thr = thr.clamp(cfg.MODEL.RUN_CONF_LOWER)
        thr_px = thr.gather(1, idx.view(B, -1)).view_as(conf)
        amb = ((conf - thr_px).abs() < 1e-5) | ((gap < 1e-5) & (conf > 0))
        out[pre + "ambiguous"] = amb.numpy()
        lab = outs["teacher_labels"]
        valid = (lab != 255).float().mean().item()
        ncls = len(torch.unique(lab[lab != 255]))
        print("step", step, {k: float(v) for k, v in losses.items()}, "valid frac %.3f" % valid,
              "classes", ncls, "ambiguous", int(amb.sum()), "logits absmax %.2f" % float(np.abs(out[pre + "logits"]).max()))
        assert 0.1 <= valid <= 0.9 and ncls >= 5, "fixture not discriminative"
        # gradients
        names, norms = [], []
        for k, p in net.backbone.named_parameters():
            names.append(k); norms.append(p.grad.double().norm().item())
        out[pre + "grad_norms"] = np.array(norms)
        if step == 0:
            out["grad_names"] = np.array(names)
        picks = ("model.conv1.weight", "model.bn1.weight", "model.bn1.bias", "model.layer1.0.conv1.weight",
                 "model.layer2.0.downsample.0.weight", "model.layer3.5.bn2.weight", "model.layer3.5.bn2.bias",
                 "model.layer3.5.conv2.weight", "model.layer4.2.conv3.weight",
                 "model.layer5.conv2d_list.1.bias", "model.layer5.conv2d_list.3.weight") if ARCH == "resnet101" else \
                ("block1.0.weight", "block1.1.bias", "block1.17.weight", "block2.27.weight", "block2.28.weight", "block3.40.bias",
                 "vgg_head.0.weight", "vgg_head.1.weight", "vgg_head.4.weight", "vgg_head.5.bias", "vgg_head.8.weight",
                 "vgg_head.8.bias", "score_pool4.weight", "score_pool3.weight", "score_pool3.bias") if ARCH == "fcn" else \
                ("features.0.weight", "features.0.bias", "features.1.weight", "features.1.bias", "features.10.weight",
                 "features.18.weight", "features.24.bias", "features.36.weight", "features.42.weight", "features.44.bias",
                 "classifier.conv2d_list.0.bias", "classifier.conv2d_list.2.weight")
        for k in picks:
            g = dict(net.backbone.named_parameters())[k].grad
            if g.numel() > 60000:
                g = g.flatten()[:60000]
            out[pre + "grad::" + k] = g.numpy().copy()
        if step == 0:
            optim.step()
            if ARCH == "resnet101":
                out["s0_post_step::model.layer5.conv2d_list.1.bias"] = \
                    net.backbone.model.layer5.conv2d_list[1].bias.detach().numpy().copy()
                out["s0_post_step::model.layer3.5.conv2.weight"] = \
                    net.backbone.model.layer3[5].conv2.weight.detach().flatten()[:60000].numpy().copy()
            elif ARCH == "vgg16":
                out["s0_post_step::features.44.bias"] = net.backbone.features[44].bias.detach().numpy().copy()
            else:
                out["s0_post_step::vgg_head.8.bias"] = net.backbone.vgg_head[8].bias.detach().numpy().copy()
    path = os.path.join(HERE, {"resnet101": "sac_resnet101_tiny.npz", "vgg16": "sac_vgg16_cfg1.npz", "fcn": "sac_fcn8s_tiny.npz"}[ARCH])
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
