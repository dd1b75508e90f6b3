"""Target-view augmentation on the GPU -- the device-side replacement of ``DataTarget.__getitem__``'s PIL pipeline
(/root/reference/datasets/dataloader_target.py:264-306, /root/reference/datasets/tf_target.py).

Split of work:
  * host (this file): draw the random parameters of every view **in the reference's order from the same generators**
    (Python ``random`` for flips / zoom-crops / blur radii / coin flips, ``torch`` for ``ColorJitter.get_params``), and
    build ``affine`` / ``affine_inv`` from them exactly like ``DataTarget._get_affine`` / ``_get_affine_inv``
    (dataloader_target.py:220-262).  With the same seeds the operators are identical to the reference's.
  * device (csrc/sacb_aug.cu, ``sacb_target_augment``): all pixel work for the G*T views in four streaming kernels.

Input is the *base crop* of each target image (uint8 RGB at crop size, i.e. the output of MaskScale / MaskRandScale /
MaskRandCrop / MaskRandHFlip, dataloader_target.py:101-106) plus its pad mask; output is the reference's
``batch_target`` already flattened to [G*T, ...]: (frames1, gt, frames2, affine, affine_inv)."""
import ctypes as C
import math
import random as _random

import torch

from . import lib as L

NPARAM = 16
MEAN = (0.485, 0.456, 0.406)            # dataloader_base.py:39-40
STD = (0.229, 0.224, 0.225)


class AugCfg(object):
    """the DATASET keys the target pipeline reads (configs/deeplabv2_resnet101_train.yaml, core/config.py:75-95)"""
    RND_ZOOM = (0.5, 1.0)
    GUIDED_HFLIP = True
    RND_BLUR = True
    RND_JITTER = 0.4
    RND_GREYSCALE = 0.2
    BLUR_RADIUS = (0.1, 2.0)            # RandGaussianBlur default (tf_target.py:337)
    JITTER_P = 0.5                      # MaskRandJitter default (tf_target.py:372)


class Aug(C.Structure):
    _fields_ = [("size", C.c_uint32), ("G", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("base", C.c_void_p), ("base_mask", C.c_void_p), ("base_label", C.c_void_p), ("view_params", C.c_void_p),
                ("mean", C.c_float * 3), ("std", C.c_float * 3),
                ("raw", C.c_void_p), ("mask", C.c_void_p), ("tmp", C.c_void_p), ("levels", C.c_void_p), ("grey_sum", C.c_void_p),
                ("frames1", C.c_void_p), ("gt", C.c_void_p), ("frames2", C.c_void_p)]


def affine_from_params(params, crop_hw):
    """DataTarget._get_affine / _get_affine_inv (dataloader_target.py:220-262); params: [(dy, dx, alpha, 1/s, flip)]"""
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
    Ai /= torch.Tensor(params)[:, 3].view(-1, 1, 1) ** 2
    return A, Ai


def draw_group_params(K, crop_hw, cfg=AugCfg, rnd=_random):
    """Random parameters of ONE view-group, consuming ``rnd`` / the global torch generator in the reference's order:
    GuidedRandHFlip (K coin flips), MaskRandScaleCrop (views 1..K-1: scale, top, left), then on copy #1 RandGaussianBlur
    (K radii), MaskRandJitter (K coin flips, each hit followed by ColorJitter.get_params on the torch generator),
    MaskRandGreyscale (K coin flips).  Returns (rows [K][16] for the kernel, affine params [K][5])."""
    H, W = crop_hw
    aff = [[0.0, 0.0, 0.0, 1.0, 1.0] for _ in range(K)]
    if cfg.GUIDED_HFLIP:                                          # tf_target.py:140-156
        for i in range(K):
            if rnd.random() > 0.5:
                aff[i][4] *= -1
    win = [(0, 0, H, W) for _ in range(K)]
    z0, z1 = cfg.RND_ZOOM
    if z1 - z0 > 0:                                               # tf_target.py:158-239
        for k in range(1, K):
            s = rnd.uniform(z0, z1)
            nh, nw = int(s * H), int(s * W)
            if s < 1.0:
                i, j = rnd.randint(0, H - nh), rnd.randint(0, W - nw)
            else:
                i, j = rnd.randint(H - nh, 0), rnd.randint(W - nw, 0)
            if s == 1.0:
                continue
            aff[k][0] = i + nh / 2 - H / 2
            aff[k][1] = j + nw / 2 - W / 2
            aff[k][3] = 1 / s
            win[k] = (i, j, nh, nw)
    sig = [0.0] * K
    if cfg.RND_BLUR:                                              # tf_target.py:331-349
        sig = [rnd.uniform(cfg.BLUR_RADIUS[0], cfg.BLUR_RADIUS[1]) for _ in range(K)]
    jit = [None] * K
    if cfg.RND_JITTER > 0:                                        # tf_target.py:366-390 + torchvision ColorJitter.get_params
        j = cfg.RND_JITTER
        lo, hi, hh = max(0.0, 1.0 - j), 1.0 + j, min(0.1, j)
        for i in range(K):
            if rnd.random() < cfg.JITTER_P:
                order = torch.randperm(4).tolist()
                b = float(torch.empty(1).uniform_(lo, hi)); c = float(torch.empty(1).uniform_(lo, hi))
                s_ = float(torch.empty(1).uniform_(lo, hi)); h = float(torch.empty(1).uniform_(-hh, hh))
                jit[i] = (order, b, c, s_, h)
    grey = [False] * K
    if cfg.RND_GREYSCALE > 0:                                     # tf_target.py:351-364
        grey = [cfg.RND_GREYSCALE > rnd.random() for _ in range(K)]
    rows = []
    for k in range(K):
        r = [0.0] * NPARAM
        r[0] = aff[k][4]
        r[1], r[2], r[3], r[4] = [float(v) for v in win[k]]
        r[5] = sig[k]
        if jit[k] is not None:
            order, b, c, s_, h = jit[k]
            r[6] = 1.0
            r[7:11] = [float(o) for o in order]
            r[11], r[12], r[13], r[14] = b, c, s_, h
        r[15] = 1.0 if grey[k] else 0.0
        rows.append(r)
    return rows, aff


class TargetAugmenter(object):
    """``aug(base_u8, base_mask=None, base_label=None)`` -> (frames1, gt, frames2, affine, affine_inv) on the device,
    flattened to [G*T, ...] like ``Trainer._prep_batch`` (train.py:186-187) hands them to ``SAC.forward``."""

    def __init__(self, group_size, crop_hw, cfg=AugCfg, mean=MEAN, std=STD, rnd=_random):
        self.K, self.hw, self.cfg, self.mean, self.std, self.rnd = group_size, tuple(crop_hw), cfg, mean, std, rnd
        self._ws = {}
        self.last_params = None

    def _workspace(self, G, dev):
        key = (G, dev)
        if key not in self._ws:
            BT, (H, W) = G * self.K, self.hw
            u8 = dict(device=dev, dtype=torch.uint8)
            self._ws[key] = dict(raw=torch.empty(BT, H, W, 3, **u8), mask=torch.empty(BT, H, W, **u8),
                                 tmp=torch.empty(BT, H, W, 3, device=dev), levels=torch.empty(BT, H, W, 3, device=dev),
                                 grey=torch.zeros(BT, device=dev, dtype=torch.int64),
                                 params=torch.empty(BT, NPARAM, device=dev))
        return self._ws[key]

    def draw(self, G):
        rows, A, Ai = [], [], []
        for _ in range(G):                                        # one DataTarget.__getitem__ per group
            r, aff = draw_group_params(self.K, self.hw, self.cfg, self.rnd)
            a, ai = affine_from_params(aff, self.hw)
            rows += r; A.append(a); Ai.append(ai)
        return torch.tensor(rows, dtype=torch.float32), torch.cat(A, 0), torch.cat(Ai, 0)

    def __call__(self, base_u8, base_mask=None, base_label=None, params=None):
        assert base_u8.dtype == torch.uint8 and base_u8.dim() == 4 and base_u8.shape[-1] == 3, \
            "base crops: uint8 [G,H,W,3] on the device"           # "on the device" is enforced by L.ptr below
        G, H, W, _ = base_u8.shape
        assert (H, W) == self.hw
        dev = base_u8.device
        rows, A, Ai = self.draw(G) if params is None else params
        self.last_params = (rows, A, Ai)
        ws = self._workspace(G, dev)
        ws["params"].copy_(rows, non_blocking=True)
        BT = G * self.K
        frames1 = torch.empty(BT, 3, H, W, device=dev)
        frames2 = torch.empty(BT, 3, H, W, device=dev)
        gt = torch.empty(BT, H, W, device=dev, dtype=torch.int64)
        f3 = C.c_float * 3
        d = Aug(C.sizeof(Aug), G, self.K, H, W, L.ptr(base_u8.contiguous()),
                L.ptr(base_mask.contiguous()) if base_mask is not None else None,
                L.ptr(base_label.contiguous()) if base_label is not None else None,
                L.ptr(ws["params"]), f3(*self.mean), f3(*self.std), L.ptr(ws["raw"]), L.ptr(ws["mask"]), L.ptr(ws["tmp"]),
                L.ptr(ws["levels"]), L.ptr(ws["grey"]), L.ptr(frames1), L.ptr(gt), L.ptr(frames2))
        L.check(L.lib().sacb_target_augment(C.byref(d), L.stream()), "sacb_target_augment")
        return frames1, gt, frames2, A.to(dev, non_blocking=True), Ai.to(dev, non_blocking=True)
