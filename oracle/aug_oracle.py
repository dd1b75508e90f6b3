"""CPU oracle for the device-side target-view augmentation -- TEST INFRASTRUCTURE ONLY (see oracle/sac_oracle.py's header:
only tests/, smoke() and bench.py's CPU legs may import this).

A numpy fp32 restatement of what /root/reference/datasets/tf_target.py asks Pillow / torchvision to do for one
view-group, written from Pillow's published 8-bit rules:
  * ``view_geometry``  -- GuidedRandHFlip (tf_target.py:140-156) + MaskRandScaleCrop (:158-239): hflip, crop (or pad),
    ``Image.resize(BILINEAR)`` = separable triangle filter with Pillow's ``precompute_coeffs`` tap rule (support scaled when
    shrinking, taps clipped to the crop, renormalised); mask / label through ``NEAREST`` (floor((o + 0.5) * scale)).
  * ``photometric``    -- RandGaussianBlur (:331-349; Pillow approximates a Gaussian of sigma = radius with 3 box passes, here
    the Gaussian itself), ColorJitter on PIL images (:366-390): ImageEnhance blends ``d + f * (img - d)`` truncated to 8
    bits with d = black / mean grey level / per-pixel grey, hue through Pillow's uint8 HSV round trip; greyscale (:351-364)
    with L = (19595 r + 38470 g + 7471 b + 0x8000) >> 16.
  * ``finish``         -- to_tensor / Normalize / ApplyMask (:32-98).
Parity pin: tests/golden/aug_reference.npz holds the outputs of the REAL reference classes (PIL) for seeded parameters
(tests/golden/make_golden_aug.py); this oracle must reproduce the affine operators / masks / labels exactly and the pixels
to about one grey level (tests/test_augment_cpu.py).  The CUDA kernels are then compared with this oracle level by level.
"""
import numpy as np

F32 = np.float32
MAXT, BLUR_R = 6, 6


def _tri(x):
    x = np.abs(x)
    return np.where(x < 1, F32(1) - x, F32(0)).astype(F32)


def resample_coeffs(out_size, scale, in_size):
    """Pillow precompute_coeffs for BILINEAR: returns kmin [out], kn [out], w [out, MAXT] (fp32)"""
    scale = F32(scale)
    fs = max(scale, F32(1))
    o = np.arange(out_size, dtype=F32)
    center = (o + F32(0.5)) * scale
    ss = F32(1) / fs
    xmin = np.maximum((center - fs + F32(0.5)).astype(np.int32), 0)
    xmax = np.minimum((center + fs + F32(0.5)).astype(np.int32), in_size)
    n = np.minimum(xmax - xmin, MAXT)
    w = np.zeros((out_size, MAXT), F32)
    ww = np.zeros(out_size, F32)
    for k in range(MAXT):
        v = np.where(k < n, _tri(((k + xmin).astype(F32) - center + F32(0.5)) * ss), F32(0)).astype(F32)
        w[:, k] = v
        ww = ww + v
    nz = ww != 0
    w[nz] = w[nz] / ww[nz, None]
    return xmin, n, w


def view_geometry(base, base_mask, base_label, row):
    """one view: base uint8 [H,W,3] -> (raw uint8 [H,W,3], mask uint8 [H,W], label uint8 [H,W])"""
    H, W, _ = base.shape
    flip = row[0] < 0
    top, left, ch, cw = int(row[1]), int(row[2]), int(row[3]), int(row[4])
    sy, sx = F32(ch) / F32(H), F32(cw) / F32(W)
    ky0, kyn, wy = resample_coeffs(H, sy, ch)
    kx0, kxn, wx = resample_coeffs(W, sx, cw)
    src = base[:, ::-1] if flip else base
    srcf = src.astype(F32)
    acc = np.zeros((H, W, 3), F32)
    for a in range(MAXT):
        yy = top + ky0 + a                                            # [H]
        vy = (a < kyn) & (yy >= 0) & (yy < H)
        rowacc = np.zeros((H, W, 3), F32)
        for c in range(MAXT):
            xx = left + kx0 + c                                       # [W]
            vx = (c < kxn) & (xx >= 0) & (xx < W)
            pix = srcf[np.clip(yy, 0, H - 1)[:, None], np.clip(xx, 0, W - 1)[None, :]]
            term = (wx[:, c][None, :, None] * pix).astype(F32)
            rowacc = np.where((vy[:, None] & vx[None, :])[..., None], rowacc + term, rowacc).astype(F32)
        acc = np.where(vy[:, None, None], acc + (wy[:, a][:, None, None] * rowacc).astype(F32), acc).astype(F32)
    raw = np.clip((acc + F32(0.5)).astype(np.int32), 0, 255).astype(np.uint8)
    ys = np.minimum(((np.arange(H, dtype=F32) + F32(0.5)) * sy).astype(np.int32), ch - 1) + top
    xs = np.minimum(((np.arange(W, dtype=F32) + F32(0.5)) * sx).astype(np.int32), cw - 1) + left
    inside = ((ys >= 0) & (ys < H))[:, None] & ((xs >= 0) & (xs < W))[None, :]
    yc, xc = np.clip(ys, 0, H - 1), np.clip(xs, 0, W - 1)
    bm = (base_mask[:, ::-1] if flip else base_mask) if base_mask is not None else np.zeros((H, W), np.uint8)
    bl = (base_label[:, ::-1] if flip else base_label) if base_label is not None else np.full((H, W), 255, np.uint8)
    mask = np.where(inside, (bm[yc[:, None], xc[None, :]] > 0).astype(np.uint8), 1).astype(np.uint8)
    label = np.where(inside, bl[yc[:, None], xc[None, :]], 255).astype(np.uint8)
    return raw, mask, label


def _clip_trunc(t):
    return np.where(t <= 0, F32(0), np.where(t >= 255, F32(255), np.floor(t))).astype(F32)


def _grey_L(v):
    r, g, b = [v[..., i].astype(np.uint32) for i in range(3)]
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(F32)


def _hue(v, hue):
    r, g, b = [v[..., i].astype(np.int32) for i in range(3)]
    maxc, minc = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
    grey = minc == maxc
    cr = np.where(grey, 1, maxc - minc).astype(F32)
    mx = np.where(maxc == 0, 1, maxc).astype(F32)
    s = cr / mx
    rc, gc, bc = (maxc - r).astype(F32) / cr, (maxc - g).astype(F32) / cr, (maxc - b).astype(F32) / cr
    h = np.where(r == maxc, bc - gc, np.where(g == maxc, F32(2) + rc - bc, F32(4) + gc - rc)).astype(F32)
    h = np.fmod((h / F32(6) + F32(1)).astype(F32), F32(1)).astype(F32)
    uh = np.where(grey, 0, np.clip((h * F32(255)).astype(np.int32), 0, 255))
    us = np.where(grey, 0, np.clip((s * F32(255)).astype(np.int32), 0, 255))
    shift = int(np.int32(F32(hue) * F32(255))) & 255
    uh = (uh + shift) & 255
    hf = (uh.astype(F32) * F32(6) / F32(255)).astype(F32)
    i = np.floor(hf).astype(np.int32)
    f = (hf - i.astype(F32)).astype(F32)
    fs = (us.astype(F32) / F32(255)).astype(F32)
    fv = maxc.astype(F32)
    one = F32(1)
    pp = np.clip(np.rint(fv * (one - fs)), 0, 255).astype(F32)
    qq = np.clip(np.rint(fv * (one - (fs * f).astype(F32))), 0, 255).astype(F32)
    tt = np.clip(np.rint(fv * (one - (fs * (one - f)).astype(F32))), 0, 255).astype(F32)
    k = i % 6
    R = np.select([k == 0, k == 1, k == 2, k == 3, k == 4], [fv, qq, pp, pp, tt], fv)
    G = np.select([k == 0, k == 1, k == 2, k == 3, k == 4], [tt, fv, fv, qq, pp], pp)
    B = np.select([k == 0, k == 1, k == 2, k == 3, k == 4], [pp, pp, tt, fv, fv], qq)
    out = np.stack([R, G, B], -1).astype(F32)
    return np.where((us == 0)[..., None], fv[..., None].repeat(3, -1), out).astype(F32)


def _blur(raw, sigma):
    H, W, _ = raw.shape
    img = raw.astype(F32)
    if not sigma > 0:
        return img
    sigma = F32(sigma)
    R = min(BLUR_R, int(np.ceil(F32(3) * sigma)))
    inv = F32(1) / (F32(2) * sigma * sigma)
    ks = np.arange(-R, R + 1)
    w = np.exp(-(ks * ks).astype(F32) * inv).astype(F32)

    def one_pass(src, axis, n):
        acc = np.zeros_like(src); ws = F32(0)
        idx = np.arange(n)
        for k, wk in zip(ks, w):
            j = np.clip(idx + k, 0, n - 1)
            acc = (acc + (wk * np.take(src, j, axis=axis)).astype(F32)).astype(F32)
            ws = F32(ws + wk)
        return (acc / ws).astype(F32)
    hpass = one_pass(img, 1, W)
    vpass = one_pass(hpass, 0, H)
    return np.clip(np.floor(vpass + F32(0.5)), 0, 255).astype(F32)


def photometric(raw, row):
    """noisy copy of one view: uint8 [H,W,3] -> fp32 grey levels [H,W,3] (integers held in floats)"""
    v = _blur(raw, row[5])
    if row[6] != 0:
        for k in range(4):
            op = int(row[7 + k])
            if op == 0:
                v = _clip_trunc(F32(0) + F32(row[11]) * (v - F32(0)))
            elif op == 1:
                mean = F32(int(float(_grey_L(v).astype(np.float64).sum()) / (v.shape[0] * v.shape[1]) + 0.5))
                v = _clip_trunc(mean + F32(row[12]) * (v - mean))
            elif op == 2:
                L = _grey_L(v)[..., None]
                v = _clip_trunc(L + F32(row[13]) * (v - L))
            else:
                v = _hue(v, row[14])
    if row[15] != 0:
        v = _grey_L(v)[..., None].repeat(3, -1).astype(F32)
    return v


def finish(levels, mask, mean, std):
    """to_tensor + Normalize + ApplyMask: levels [H,W,3] -> fp32 [3,H,W]"""
    t = (levels.astype(F32) / F32(255) - np.asarray(mean, F32)) / np.asarray(std, F32)
    t = np.where(mask[..., None] > 0, F32(0), t).astype(F32)
    return np.ascontiguousarray(t.transpose(2, 0, 1))


def augment_group(base, base_mask, base_label, rows, mean, std):
    """one view-group -> (frames1 [K,3,H,W], gt int64 [K,H,W], frames2 [K,3,H,W], levels1 [K,H,W,3], raw [K,H,W,3])"""
    f1, gt, f2, lv, rw = [], [], [], [], []
    for row in rows:
        raw, mask, label = view_geometry(base, base_mask, base_label, row)
        lev = photometric(raw, row)
        f2.append(finish(raw.astype(F32), mask, mean, std))
        f1.append(finish(lev, mask, mean, std))
        gt.append(np.where(mask > 0, -1, label.astype(np.int64)))
        lv.append(lev); rw.append(raw)
    return np.stack(f1), np.stack(gt), np.stack(f2), np.stack(lv), np.stack(rw)
