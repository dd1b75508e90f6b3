"""Seeded synthetic weights and target batches for the SAC target step.

Nothing here touches the GPU path; it only manufactures the tensors that the
reference's DataLoader / checkpoint would have supplied:

* ``make_backbone_params`` -- a state_dict for ``DeepLabV2_ResNet101``
  (key layout of /root/reference/models/deeplabv2.py:118-171) with the seeded
  He-fan-in recipe of SURVEY.md section 8(d).  The reference's stock
  ``normal_(0, 0.01)`` init (deeplabv2.py:135-141) collapses activations and
  yields all-255 pseudo-label masks, which would make parity vacuous.
* ``make_target_batch`` -- ``(frames1, y, frames2, affine, affine_inv)`` in the
  convention of ``DataTarget.__getitem__`` (/root/reference/datasets/
  dataloader_target.py:264-306): view 0 is the un-zoomed original, views k>=1
  are zoom-crops; ``affine`` / ``affine_inv`` follow ``_get_affine`` /
  ``_get_affine_inv`` (dataloader_target.py:220-262).
"""
from collections import OrderedDict
import math

import torch
import torch.nn.functional as F

RESNET101_LAYERS = (3, 4, 23, 3)
NUM_CLASSES = 19


def resnet101_conv_table(num_classes=NUM_CLASSES):
    """(key prefix, Cout, Cin, k, stride, dilation, pad, has_bn, has_bias) for
    every conv of DeepLabV2_ResNet101 in state_dict order
    (/root/reference/models/deeplabv2.py:54-171)."""
    t = [("model.conv1", "model.bn1", 64, 3, 7, 2, 1, 3)]
    inplanes = 64
    cfg = ((64, RESNET101_LAYERS[0], 1, 1), (128, RESNET101_LAYERS[1], 2, 1),
           (256, RESNET101_LAYERS[2], 1, 2), (512, RESNET101_LAYERS[3], 1, 4))
    for li, (planes, blocks, stride, dil) in enumerate(cfg, start=1):
        for b in range(blocks):
            p = "model.layer%d.%d" % (li, b)
            s = stride if b == 0 else 1
            t.append((p + ".conv1", p + ".bn1", planes, inplanes, 1, s, 1, 0))
            t.append((p + ".conv2", p + ".bn2", planes, planes, 3, 1, dil, dil))
            t.append((p + ".conv3", p + ".bn3", planes * 4, planes, 1, 1, 1, 0))
            if b == 0:
                t.append((p + ".downsample.0", p + ".downsample.1", planes * 4, inplanes, 1, s, 1, 0))
            inplanes = planes * 4
    for i, d in enumerate((6, 12, 18, 24)):
        t.append(("model.layer5.conv2d_list.%d" % i, None, num_classes, 2048, 3, 1, d, d))
    return t


def make_backbone_params(seed=123, num_classes=NUM_CLASSES, dtype=torch.float32):
    """Seeded ResNet-101 DeepLabv2 state_dict (632 entries incl.
    num_batches_tracked), iterated in the reference's state_dict order."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def bn(prefix, c, gamma_scale=1.0):
        sd[prefix + ".weight"] = (torch.rand(c, generator=g) * 0.4 + 0.8) * gamma_scale
        sd[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_var"] = torch.rand(c, generator=g) * 0.4 + 0.8
        sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    # state_dict order of the reference module tree: conv, bn pairs; in a
    # Bottleneck: conv1,bn1,conv2,bn2,conv3,bn3,downsample.0,downsample.1
    for (ck, bk, cout, cin, k, s, d, p) in resnet101_conv_table(num_classes):
        fan_in = cin * k * k
        sd[ck + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / fan_in)
        if bk is None:
            sd[ck + ".bias"] = torch.zeros(cout)
        else:
            bn(bk, cout, 0.2 if bk.endswith("bn3") else 1.0)
    return OrderedDict((k, v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items())


VGG16_CONVS = ((0, 3, 64, 1), (3, 64, 64, 1), (7, 64, 128, 1), (10, 128, 128, 1), (14, 128, 256, 1), (17, 256, 256, 1),
               (20, 256, 256, 1), (24, 256, 512, 1), (27, 512, 512, 1), (30, 512, 512, 1), (33, 512, 512, 2),
               (36, 512, 512, 2), (39, 512, 512, 2))


def make_vgg16_params(seed=321, num_classes=NUM_CLASSES):
    """Seeded DeepLabV2_VGG16(use_bn=True) state_dict (key layout of /root/reference/models/deeplabv2.py:229-312):
    He fan-in conv weights, small random biases, BN gamma~U(0.8,1.2), beta/mean~N(0,0.1), var~U(0.8,1.2),
    ASPP weights x3 so that the pseudo labels are discriminative."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for (idx, cin, cout, dil) in VGG16_CONVS:
        sd["features.%d.weight" % idx] = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))
        sd["features.%d.bias" % idx] = torch.randn(cout, generator=g) * 0.05
        b = "features.%d" % (idx + 1)
        sd[b + ".weight"] = torch.rand(cout, generator=g) * 0.4 + 0.8
        sd[b + ".bias"] = torch.randn(cout, generator=g) * 0.1
        sd[b + ".running_mean"] = torch.randn(cout, generator=g) * 0.1
        sd[b + ".running_var"] = torch.rand(cout, generator=g) * 0.4 + 0.8
        sd[b + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    for idx, cin, cout in ((42, 512, 1024), (44, 1024, 1024)):
        sd["features.%d.weight" % idx] = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))
        sd["features.%d.bias" % idx] = torch.randn(cout, generator=g) * 0.05
    for i in range(4):
        sd["classifier.conv2d_list.%d.weight" % i] = torch.randn(num_classes, 1024, 3, 3, generator=g) * math.sqrt(2.0 / (1024 * 9)) * 3.0
        sd["classifier.conv2d_list.%d.bias" % i] = torch.randn(num_classes, generator=g) * 0.05
    return sd


def affine_from_params(params, crop_hw):
    """Restates DataTarget._get_affine/_get_affine_inv
    (/root/reference/datasets/dataloader_target.py:220-262).
    params: list of (dy, dx, alpha_deg, scale, flip)."""
    H, W = crop_hw
    ar = float(H) / float(W)
    K = len(params)
    A = torch.zeros(K, 2, 3)
    for i, (dy, dx, alpha, scale, flip) in enumerate(params):
        sin = math.sin(alpha * math.pi / 180.0)
        cos = math.cos(alpha * math.pi / 180.0)
        A[i, 0, 0], A[i, 0, 1] = flip * cos, sin * ar
        A[i, 1, 0], A[i, 1, 1] = -sin / ar, cos
        A[i, 0, 2] = -1.0 * (cos * dx + sin * dy)
        A[i, 1, 2] = -1.0 * (-sin * dx + cos * dy)
        A[i, 0, 2] /= float(W // 2)
        A[i, 1, 2] /= float(H // 2)
        A[i] *= scale
    Ai = A.clone()
    Ai[:, 0, 1] = A[:, 1, 0] * ar ** 2
    Ai[:, 1, 0] = A[:, 0, 1] / ar ** 2
    Ai[:, 0, 2] = -1 * (Ai[:, 0, 0] * A[:, 0, 2] + Ai[:, 0, 1] * A[:, 1, 2])
    Ai[:, 1, 2] = -1 * (Ai[:, 1, 0] * A[:, 0, 2] + Ai[:, 1, 1] * A[:, 1, 2])
    Ai /= torch.tensor([p[3] for p in params], dtype=torch.float32).view(-1, 1, 1) ** 2
    return A, Ai


def make_target_batch(num_groups, group_size, crop_hw, seed=0, ignore_rows=8):
    """Synthetic ``batch_target`` already flattened to [B*T, ...] like
    Trainer._prep_batch (/root/reference/train.py:157-187) returns it.

    frames2: per group a smooth-noise base image; view 0 = identity (random
    flip), views k>=1 = zoom-crop s~U(0.5,1) with a uniform offset and random
    flip, resampled from the base.  frames1 = frames2 with per-channel gain and
    additive noise (stand-in for blur/jitter/greyscale, tf_target.py:331-390).
    y = 255 everywhere except ``ignore_rows`` bottom rows of the zoomed views
    set to -1 (the augmentation-padding marker, tf_target.py:84-98)."""
    H, W = crop_hw
    g = torch.Generator().manual_seed(seed)

    def U(a, b):
        return float(torch.rand((), generator=g)) * (b - a) + a

    f1, f2, ys, As, Ais = [], [], [], [], []
    for _ in range(num_groups):
        base = torch.randn(1, 3, max(H // 8, 2), max(W // 8, 2), generator=g)
        base = F.interpolate(base, size=(H, W), mode="bicubic", align_corners=False)
        base = base + 0.1 * torch.randn(1, 3, H, W, generator=g)
        params = []
        for k in range(group_size):
            flip = 1.0 if U(0, 1) > 0.5 else -1.0
            if k == 0:
                params.append((0.0, 0.0, 0.0, 1.0, flip))
            else:
                s = U(0.5, 1.0)
                dx = U(-(1 - s), (1 - s)) * (W // 2)
                dy = U(-(1 - s), (1 - s)) * (H // 2)
                params.append((dy, dx, 0.0, 1.0 / s, flip))
        A, Ai = affine_from_params(params, crop_hw)
        grid = F.affine_grid(Ai, size=(group_size, 3, H, W), align_corners=False)
        views = F.grid_sample(base.expand(group_size, -1, -1, -1), grid, mode="bilinear",
                              padding_mode="zeros", align_corners=False)
        gain = torch.rand(group_size, 3, 1, 1, generator=g) * 0.8 + 0.6
        noisy = views * gain + 0.05 * torch.randn(group_size, 3, H, W, generator=g)
        y = torch.full((group_size, H, W), 255, dtype=torch.long)
        # a labelled block so that the monitoring loss_ce (deeplabv2.py:223-224) is exercised
        cells = torch.randint(0, NUM_CLASSES, (group_size, max(H // 32, 1), max(W // 32, 1)), generator=g)
        blk = cells.repeat_interleave(16, 1).repeat_interleave(16, 2)
        y[:, :blk.shape[1], :blk.shape[2]] = blk
        if ignore_rows > 0:
            y[1:, H - ignore_rows:, :] = -1
        f1.append(noisy); f2.append(views); ys.append(y); As.append(A); Ais.append(Ai)
    cat = lambda xs: torch.cat(xs, 0).contiguous()
    return cat(f1), cat(ys), cat(f2), cat(As), cat(Ais)


def make_source_batch(n, crop_hw, seed=0, ignore_rows=8):
    """``(image, mask)`` of the source loader (what ``Trainer.step`` consumes, /root/reference/train.py:119-125):
    ``image`` fp32 [n,3,H,W] (smooth noise, like the target frames), ``mask`` int64 [n,H,W] with classes 0..18 in 16x16
    cells and 255 (ignore) in the last ``ignore_rows`` rows."""
    H, W = crop_hw
    g = torch.Generator().manual_seed(1000 + seed)
    base = torch.randn(n, 3, max(H // 8, 2), max(W // 8, 2), generator=g)
    img = F.interpolate(base, size=(H, W), mode="bicubic", align_corners=False) + 0.1 * torch.randn(n, 3, H, W, generator=g)
    cells = torch.randint(0, NUM_CLASSES, (n, (H + 15) // 16, (W + 15) // 16), generator=g)
    y = cells.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :H, :W].contiguous()
    if ignore_rows > 0:
        y[:, H - ignore_rows:, :] = 255
    return img.contiguous(), y


class ModelCfg(object):
    """Hot-path keys of cfg.MODEL (/root/reference/core/config.py:134-159)
    with the values of configs/deeplabv2_resnet101_train.yaml:22-33."""
    ARCH = "deeplabv2_resnet101"
    INIT_MODEL = ""
    BASELINE = False
    LR = 2.5e-4
    LR_TARGET = 5.0
    MOMENTUM = 0.9
    WEIGHT_DECAY = 5e-4
    STAT_MOMENTUM = 0.99
    NET_MOMENTUM = 0.99
    NET_MOMENTUM_ITER = 100
    CONF_DISCOUNT = True
    CONF_POOL_ON = True
    CONF_POOL = "avg_pool"
    FOCAL_P = 3
    LOSS = "focal_ce_conf"
    RUN_CONF_UPPER = 0.75
    RUN_CONF_LOWER = 0.2
    THRESHOLD_BETA = 1e-3
    OPT = "SGD"
    OPT_NESTEROV = False


class ModelCfgVGG16(ModelCfg):
    """cfg.MODEL of configs/deeplabv2_vgg16_train.yaml:22-33"""
    ARCH = "deeplabv2_vgg16_bn"
    LR_TARGET = 2.0
    RUN_CONF_LOWER = 0.1


FCN_CONVS = (("block1", (0, 3, 64)), ("block1", (3, 64, 64)), ("block1", (7, 64, 128)), ("block1", (10, 128, 128)),
             ("block1", (14, 128, 256)), ("block1", (17, 256, 256)), ("block1", (20, 256, 256)),
             ("block2", (24, 256, 512)), ("block2", (27, 512, 512)), ("block2", (30, 512, 512)),
             ("block3", (34, 512, 512)), ("block3", (37, 512, 512)), ("block3", (40, 512, 512)))


def make_fcn_params(seed=213, num_classes=NUM_CLASSES):
    """Seeded VGG16_FCN8s(use_bn=True) state_dict (key layout of /root/reference/models/fcn.py:23-98)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def bn(prefix, c):
        sd[prefix + ".weight"] = torch.rand(c, generator=g) * 0.4 + 0.8
        sd[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_var"] = torch.rand(c, generator=g) * 0.4 + 0.8
        sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def conv(name, cout, cin, k, gain=1.0):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k)) * gain
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.05

    for blk, (idx, cin, cout) in FCN_CONVS:
        conv("%s.%d" % (blk, idx), cout, cin, 3)
        bn("%s.%d" % (blk, idx + 1), cout)
    conv("vgg_head.0", 4096, 512, 7); bn("vgg_head.1", 4096)
    conv("vgg_head.4", 4096, 4096, 1); bn("vgg_head.5", 4096)
    conv("vgg_head.8", num_classes, 4096, 1, gain=4.0)
    conv("score_pool4", num_classes, 512, 1, gain=2.0)
    conv("score_pool3", num_classes, 256, 1, gain=2.0)
    return sd


class ModelCfgFCN(ModelCfg):
    """cfg.MODEL of configs/fcn_vgg16_train.yaml"""
    ARCH = "fcn_vgg16_bn"
    LR = 5e-4
    LR_TARGET = 2.0
    RUN_CONF_LOWER = 0.1
