"""CPU oracle for the SAC target step -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this module.  The product path
(``da_sac_b200``) never does: it fails loudly when ``libsac_b200.so`` is missing.

What it is: a functional PyTorch-CPU fp32 restatement of the reference's hot
path (``Trainer._step_target`` -> ``SAC.forward`` -> backward), written from the
reference's behaviour, each function citing the file:line it follows
(paths relative to /root/reference).

Parity pin: the reference ships no tests / golden vectors,
so the pin is generated: ``tests/golden/make_golden.py`` imports the *real*
reference modules from /root/reference in the build container, runs them on the
seeded weights/inputs of ``da_sac_b200.synth`` and stores the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against
those files.  The arithmetic itself lives in PyTorch (torch 2.11.0+cu128,
torchvision 0.26.0 as installed -- the reference does not pin a version):
``F.interpolate(bilinear, align_corners=True)``, ``F.softmax``,
``F.affine_grid`` + ``F.grid_sample(bilinear, zeros, align_corners=False)``,
``F.cross_entropy``; ``upsample_explicit`` / ``warp_explicit`` below restate
those published formulas without calling ATen's kernels and are cross-checked
in the tests.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.SyncBatchNorm default eps (deeplabv2.py:14)


# --------------------------------------------------------------------------
# Backbone: DeepLabV2_ResNet101 (models/deeplabv2.py:54-227)
# --------------------------------------------------------------------------

def _bn_eval(x, p, prefix):
    # frozen BN: eval-mode statistics, trainable affine (basenet.py:86-100)
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"],
                        p[prefix + ".weight"], p[prefix + ".bias"], False, 0.0, BN_EPS)


BN_MOMENTUM = 0.1  # nn.SyncBatchNorm default momentum (deeplabv2.py:15,28: no momentum argument)


def bn_train_recorder(new_stats, momentum=BN_MOMENTUM, taps=None):
    """Training-mode BN of the ABN baseline (models/__init__.py:29: freeze_bn = not cfg.BASELINE, so BaseNet.train()
    leaves every SyncBatchNorm in training mode, basenet.py:86-100).  Without an initialised process group
    nn.SyncBatchNorm.forward falls back to F.batch_norm(..., training=True, momentum, eps): batch statistics
    (biased variance) normalise, the running statistics move by ``momentum`` (unbiased variance).  Returns a
    ``bn(x, p, prefix)`` callable that leaves ``p`` untouched and records the updated running statistics in
    ``new_stats`` (functional form of the in-place buffer update)."""
    def bn(x, p, prefix):
        rm = p[prefix + ".running_mean"].detach().clone()
        rv = p[prefix + ".running_var"].detach().clone()
        if taps is not None:
            taps[prefix + ".z"] = x
        out = F.batch_norm(x, rm, rv, p[prefix + ".weight"], p[prefix + ".bias"], True, momentum, BN_EPS)
        new_stats[prefix + ".running_mean"] = rm
        new_stats[prefix + ".running_var"] = rv
        return out
    return bn


def _bottleneck(x, p, prefix, stride, dilation, has_ds, bn=_bn_eval):
    # Bottleneck.forward (deeplabv2.py:77-99); stride sits on conv1 (:59)
    out = F.conv2d(x, p[prefix + ".conv1.weight"], None, stride)
    out = F.relu(bn(out, p, prefix + ".bn1"))
    out = F.conv2d(out, p[prefix + ".conv2.weight"], None, 1, dilation, dilation)
    out = F.relu(bn(out, p, prefix + ".bn2"))
    out = F.conv2d(out, p[prefix + ".conv3.weight"], None, 1)
    out = bn(out, p, prefix + ".bn3")
    if has_ds:
        res = F.conv2d(x, p[prefix + ".downsample.0.weight"], None, stride)
        res = bn(res, p, prefix + ".downsample.1")
    else:
        res = x
    return F.relu(out + res)


def resnet101_logits(p, x, taps=None, bn=_bn_eval):
    """ResNet.forward (deeplabv2.py:160-171). ``p``: state_dict-like mapping with
    keys ``model.*``. Returns logits [n,19,h,w]. ``taps`` (optional dict) gets
    intermediate activations for layer-wise kernel tests.  ``bn``: ``_bn_eval`` (frozen BN, the SAC path) or a
    ``bn_train_recorder`` (ABN baseline)."""
    x = F.conv2d(x, p["model.conv1.weight"], None, 2, 3)
    x = F.relu(bn(x, p, "model.bn1"))
    if taps is not None: taps["stem"] = x
    x = F.max_pool2d(x, 3, 2, 1, ceil_mode=True)          # deeplabv2.py:126
    if taps is not None: taps["pool"] = x
    cfg = ((64, 3, 1, 1), (128, 4, 2, 1), (256, 23, 1, 2), (512, 3, 1, 4))
    for li, (planes, blocks, stride, dil) in enumerate(cfg, start=1):
        for b in range(blocks):
            x = _bottleneck(x, p, "model.layer%d.%d" % (li, b), stride if b == 0 else 1, dil, b == 0, bn)
        if taps is not None: taps["layer%d" % li] = x
    # Classifier_Module.forward (deeplabv2.py:112-116): sum of 4 dilated convs
    out = None
    for i, d in enumerate((6, 12, 18, 24)):
        o = F.conv2d(x, p["model.layer5.conv2d_list.%d.weight" % i],
                     p["model.layer5.conv2d_list.%d.bias" % i], 1, d, d)
        out = o if out is None else out + o
    return out


VGG16_CONVS = ((0, 1), (3, 1), (7, 1), (10, 1), (14, 1), (17, 1), (20, 1), (24, 1), (27, 1), (30, 1), (33, 2), (36, 2), (39, 2))
VGG16_POOL_AFTER = (3, 10, 20)


def vgg16_deeplab_logits(p, x, taps=None, bn=_bn_eval):
    """DeepLabV2_VGG16._backbone (deeplabv2.py:294-298): vgg16_bn features (conv5 dilation 2, pool4/5 removed,
    :238-260), fc6/fc7 3x3 dilation 4 + ReLU (:262-265), Classifier_Module on 1024 channels (:270)."""
    for idx, dil in VGG16_CONVS:
        x = F.conv2d(x, p["features.%d.weight" % idx], p["features.%d.bias" % idx], 1, dil, dil)
        x = F.relu(bn(x, p, "features.%d" % (idx + 1)))
        if taps is not None: taps["features.%d" % idx] = x
        if idx in VGG16_POOL_AFTER:
            x = F.max_pool2d(x, 2, 2)
    for idx in (42, 44):
        x = F.relu(F.conv2d(x, p["features.%d.weight" % idx], p["features.%d.bias" % idx], 1, 4, 4))
        if taps is not None: taps["features.%d" % idx] = x
    out = None
    for i, d in enumerate((6, 12, 18, 24)):
        o = F.conv2d(x, p["classifier.conv2d_list.%d.weight" % i], p["classifier.conv2d_list.%d.bias" % i], 1, d, d)
        out = o if out is None else out + o
    return out


FCN_TRUNK = (("block1", (0, 3, 7, 10, 14, 17, 20), (3, 10, 20)), ("block2", (24, 27, 30), (30,)), ("block3", (34, 37, 40), (40,)))


def vgg16_fcn8s_logits(p, x, bn=_bn_eval):
    """VGG16_FCN8s._backbone (fcn.py:111-137) with drop_rate = 0 (Dropout2d is the identity): pool3/4/5 features, 7x7 /
    1x1 head, score_pool4 / score_pool3, bilinear x2 (align_corners=True) fusion -> scores at 1/8 resolution."""
    feats = {}
    for blk, convs, pools in FCN_TRUNK:
        for idx in convs:
            x = F.conv2d(x, p["%s.%d.weight" % (blk, idx)], p["%s.%d.bias" % (blk, idx)], 1, 1)
            x = F.relu(bn(x, p, "%s.%d" % (blk, idx + 1)))
            if idx in pools:
                x = F.max_pool2d(x, 2, 2)
        feats[blk] = x
    h = F.relu(bn(F.conv2d(feats["block3"], p["vgg_head.0.weight"], p["vgg_head.0.bias"], 1, 3), p, "vgg_head.1"))
    h = F.relu(bn(F.conv2d(h, p["vgg_head.4.weight"], p["vgg_head.4.bias"]), p, "vgg_head.5"))
    score = F.conv2d(h, p["vgg_head.8.weight"], p["vgg_head.8.bias"])
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)       # fcn.py:107-109
    score = up(score) + F.conv2d(feats["block2"], p["score_pool4.weight"], p["score_pool4.bias"])
    score = up(score) + F.conv2d(feats["block1"], p["score_pool3.weight"], p["score_pool3.bias"])
    return score


def backbone_forward(p, im, y=None):
    """DeepLabV2_ResNet101.forward / DeepLabV2_VGG16.forward (deeplabv2.py:213-227, 300-312); the architecture is
    recognised from the state_dict keys."""
    if "model.conv1.weight" in p: logits = resnet101_logits(p, im)
    elif "vgg_head.0.weight" in p: logits = vgg16_fcn8s_logits(p, im)
    else: logits = vgg16_deeplab_logits(p, im)
    logits_up = F.interpolate(logits, im.shape[-2:], mode="bilinear", align_corners=True)
    if y is None:
        return logits, logits_up
    ce = F.cross_entropy(logits_up, y, ignore_index=255, reduction="none")
    return {"loss_ce": ce.mean().view(1)}, {"logits_up": logits_up, "logits": logits}


# --------------------------------------------------------------------------
# Explicit restatements of the ATen formulas
# --------------------------------------------------------------------------

def upsample_explicit(X, H, W):
    """bilinear, align_corners=True (aten upsample_bilinear2d):
    sy = i*(h-1)/(H-1); y0=floor(sy); y1=min(y0+1,h-1); ly=sy-y0."""
    n, c, h, w = X.shape
    def axis(o, i):
        scale = (i - 1) / (o - 1) if o > 1 else 0.0
        s = torch.arange(o, dtype=torch.float32) * torch.tensor(scale, dtype=torch.float32)
        i0 = s.floor().long().clamp(max=i - 1)
        i1 = (i0 + 1).clamp(max=i - 1)
        l1 = s - i0.float()
        return i0, i1, l1
    y0, y1, ly = axis(H, h)
    x0, x1, lx = axis(W, w)
    ly = ly.view(1, 1, H, 1); lx = lx.view(1, 1, 1, W)
    top = X[:, :, y0][:, :, :, x0] * (1 - lx) + X[:, :, y0][:, :, :, x1] * lx
    bot = X[:, :, y1][:, :, :, x0] * (1 - lx) + X[:, :, y1][:, :, :, x1] * lx
    return top * (1 - ly) + bot * ly


def warp_explicit(X, M):
    """grid_sample(X, affine_grid(M, align_corners=False), bilinear, zeros,
    align_corners=False): x=(2j+1)/W-1, (u,v)=M.(x,y,1), ix=((u+1)W-1)/2,
    4 taps, out-of-range taps contribute 0 (models/sac.py:289-290,300-301,309-310)."""
    n, c, H, W = X.shape
    xs = (2 * torch.arange(W, dtype=torch.float32) + 1) / W - 1
    ys = (2 * torch.arange(H, dtype=torch.float32) + 1) / H - 1
    gx = xs.view(1, 1, W).expand(n, H, W)
    gy = ys.view(1, H, 1).expand(n, H, W)
    M = M.float()
    u = M[:, 0, 0].view(n, 1, 1) * gx + M[:, 0, 1].view(n, 1, 1) * gy + M[:, 0, 2].view(n, 1, 1)
    v = M[:, 1, 0].view(n, 1, 1) * gx + M[:, 1, 1].view(n, 1, 1) * gy + M[:, 1, 2].view(n, 1, 1)
    ix = ((u + 1) * W - 1) / 2
    iy = ((v + 1) * H - 1) / 2
    x0 = ix.floor(); y0 = iy.floor(); x1 = x0 + 1; y1 = y0 + 1
    out = torch.zeros_like(X)
    flat = X.reshape(n, c, H * W)
    for (xx, yy, wgt) in ((x0, y0, (x1 - ix) * (y1 - iy)), (x1, y0, (ix - x0) * (y1 - iy)),
                          (x0, y1, (x1 - ix) * (iy - y0)), (x1, y1, (ix - x0) * (iy - y0))):
        ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long().view(n, 1, H * W).expand(n, c, H * W)
        val = flat.gather(2, idx).view(n, c, H, W)
        out = out + val * (wgt * ok.float()).view(n, 1, H, W)
    return out


def _warp(X, M):
    grid = F.affine_grid(M, size=list(X.shape), align_corners=False)
    return F.grid_sample(X, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


# --------------------------------------------------------------------------
# SAC tail (models/sac.py)
# --------------------------------------------------------------------------

def update_running_conf(running_conf, probs, cfg, tolerance=1e-8):
    """SAC._update_running_conf (sac.py:104-117); returns the new buffer."""
    B, C, H, W = probs.shape
    probs_avg = probs.mean(0).view(C, -1).mean(-1)
    rc = running_conf.clone()
    new_index = (probs_avg > tolerance) & (rc == cfg.THRESHOLD_BETA)
    rc[new_index] = probs_avg[new_index]
    rc = rc * cfg.STAT_MOMENTUM
    rc = rc + (1 - cfg.STAT_MOMENTUM) * probs_avg
    return rc


def avg_pool(probs, T, tolerance=0.1):
    """SAC._avg_pool (sac.py:238-269), world_size 1 (no _gather)."""
    _, C, H, W = probs.shape
    probs_T = probs.view(-1, T, C, H, W)
    s = probs_T.sum(1, keepdim=True)
    z = s.sum(2, keepdim=True)
    mask = (z > tolerance).type_as(probs)
    avg = s / z.clamp(1e-3)
    avg = avg.expand(-1, T, -1, -1, -1)
    mask = mask.expand(-1, T, -1, -1, -1)
    return avg.flatten(0, 1), mask.flatten(0, 1)


def entropy(probs, eps=1e-5):
    """SAC._entropy (sac.py:189-196)"""
    probs_eps = (probs + eps) / (1 + eps)
    ent = -(probs * torch.log(probs_eps)).sum(1, keepdim=True)
    ent[probs.sum(1, keepdim=True) < 0.1] = 1.0 / eps
    return ent


def minentropy_pool(probs, T, tolerance=0.1):
    """SAC._minentropy_pool (sac.py:218-236): every view of a group receives the distribution of the group's
    lowest-entropy view (first view on ties); mask = total mass over views and classes > tolerance."""
    BT, C, H, W = probs.shape
    ent_T = entropy(probs).view(-1, T, 1, H, W)
    sel = ent_T.argmin(1, keepdim=True).expand(-1, -1, C, -1, -1)
    probs_T = probs.view(-1, T, C, H, W)
    masks = probs_T.sum(1, keepdim=True).sum(2, keepdim=True) > tolerance
    picked = probs_T.gather(1, sel).expand(-1, T, -1, -1, -1)
    masks = masks.expand(-1, T, -1, -1, -1).type_as(probs)
    return picked.reshape(BT, C, H, W), masks.reshape(BT, 1, H, W)


def refine(slow_logits, hw, T, affine, affine_inv, ignore_mask, running_conf, cfg, training=True, pool="avg_pool"):
    """SAC._refine (sac.py:271-313). Returns (teacher_refined, new running_conf, diags).
    ``pool``: "avg_pool" (default), "minentropy_pool", or None for MODEL.CONF_POOL_ON = False (sac.py:284-285)."""
    H, W = hw
    up = F.interpolate(slow_logits, (H, W), mode="bilinear", align_corners=True)   # :275
    probs = F.softmax(up, 1)                                                       # :276
    if training:
        running_conf = update_running_conf(running_conf, probs, cfg)               # :278-279
    probs = probs * (1 - ignore_mask[:, None].type_as(probs))                      # :282
    if pool is None:
        return probs, running_conf, {"teacher_init": up}
    aligned = _warp(probs, affine)                                                 # :289-290
    valid_aligned = _warp(torch.ones_like(aligned), affine_inv)                    # :299-301
    pool_fn = {"avg_pool": avg_pool, "minentropy_pool": minentropy_pool}[pool]
    refined_aligned, valid = pool_fn(aligned * valid_aligned, T)                   # :305
    grid_inv = F.affine_grid(affine_inv, size=list(probs.shape), align_corners=False)
    refined = F.grid_sample(refined_aligned, grid_inv, align_corners=False)        # :309
    refined_valid = F.grid_sample(valid, grid_inv, align_corners=False)            # :310
    refined = refined * refined_valid                                              # :311
    return refined, running_conf, {"teacher_aligned": aligned, "teacher_init": up}


def refine_fractional(slow_logits_parts, hw, T, affine_parts, affine_inv_parts, ignore_parts, running_conf_parts, cfg,
                      training=True):
    """SAC._refine when ONE view-group is spread over several ranks (train.py:185-209; sac.py:198-216,243-245):
    element r of every ``*_parts`` list is what rank r holds (T0 = T / len(parts) views).  ``_gather`` concatenates the
    ranks' ``aligned * valid`` tensors in rank order, ``_avg_pool`` pools over all T views and hands every rank T0 copies
    of the pooled map, which each rank warps back with its own ``affine_inv``.  Returns [(refined_r, running_conf_r)]."""
    H, W = hw
    prods, rcs = [], []
    for lg, A, Ai, ign, rc in zip(slow_logits_parts, affine_parts, affine_inv_parts, ignore_parts, running_conf_parts):
        up = F.interpolate(lg, (H, W), mode="bilinear", align_corners=True)
        probs = F.softmax(up, 1)
        if training:
            rc = update_running_conf(rc, probs, cfg)          # local statistics of each rank (sac.py:278-279)
        rcs.append(rc)
        probs = probs * (1 - ign[:, None].type_as(probs))
        aligned = _warp(probs, A)
        valid_aligned = _warp(torch.ones_like(aligned), Ai)
        prods.append(aligned * valid_aligned)
    gathered = torch.cat(prods, 0)                            # _gather (sac.py:214)
    assert gathered.shape[0] == T
    pooled, valid = avg_pool(gathered, T)
    out = []
    for r, Ai in enumerate(affine_inv_parts):
        T0 = Ai.shape[0]
        grid_inv = F.affine_grid(Ai, size=[T0] + list(pooled.shape[1:]), align_corners=False)
        refined = F.grid_sample(pooled[:T0], grid_inv, align_corners=False)
        refined_valid = F.grid_sample(valid[:T0], grid_inv, align_corners=False)
        out.append((refined * refined_valid, rcs[r]))
    return out


def pseudo_labels_probs(probs, ignore_augm, running_conf, cfg, discount=True):
    """SAC._pseudo_labels_probs + _threshold_discount (sac.py:151-187)."""
    B, C, H, W = probs.shape
    max_conf, max_idx = probs.max(1, keepdim=True)
    peaks = torch.zeros_like(probs)
    peaks.scatter_(1, max_idx, max_conf)
    top_peaks, _ = peaks.view(B, C, -1).max(-1)
    top_peaks = top_peaks * cfg.RUN_CONF_UPPER
    if discount:
        top_peaks = top_peaks * (1.0 - torch.exp(-running_conf / cfg.THRESHOLD_BETA)).view(1, C)
    top_peaks = top_peaks.clamp(cfg.RUN_CONF_LOWER)
    above = peaks.gt(top_peaks.view(B, C, 1, 1)).type_as(probs)
    ignore = above.sum(1, keepdim=True) != 1
    labels = max_idx.clone()
    labels[ignore] = 255
    labels = labels.squeeze(1)
    labels[ignore_augm] = 255
    return labels, max_conf, max_idx, top_peaks


def focal_ce_conf(logits_up, pseudo_gt, teacher_conf, running_conf, p=3):
    """SAC._focal_ce_conf (sac.py:134-149) INCLUDING the [B,H,W]*[B,1,H,W]
    -> [B,B,H,W] broadcast of :148 (written here as the equivalent
    mean_{j,p} ce[j,p] * mean_i conf[i,p])."""
    w = (1 - running_conf.clamp(0.0)) ** p
    ce = F.cross_entropy(logits_up, pseudo_gt, weight=w, ignore_index=255, reduction="none")
    m = teacher_conf.mean(0)              # [1,H,W] batch-mean confidence
    return (ce * m).mean()


def focal_ce(logits_up, pseudo_gt, running_conf, p=3):
    """SAC._focal_ce (sac.py:119-132): MODEL.LOSS = focal_ce, no confidence weighting."""
    w = (1 - running_conf.clamp(0.0)) ** p
    return F.cross_entropy(logits_up, pseudo_gt, weight=w, ignore_index=255, reduction="none").mean()


def focal_ce_conf_literal(logits_up, pseudo_gt, teacher_conf, running_conf, p=3):
    """Literal form of sac.py:148 (materialises [B,B,H,W]); small inputs only."""
    w = (1 - running_conf.clamp(0.0)) ** p
    ce = F.cross_entropy(logits_up, pseudo_gt, weight=w, ignore_index=255, reduction="none")
    return (ce * teacher_conf).mean()


def teacher_diff(slow, fast):
    """distance part of SAC._momentum_update (sac.py:83-102)."""
    d = torch.zeros(())
    for k, v in fast.items():
        if k.split(".")[-1] in ("weight", "bias", "running_mean", "running_var"):
            d = d + torch.norm(slow[k] - v.detach())
    return d.view(1)


def momentum_update(slow, fast, m):
    """EMA part of SAC._momentum_update with update=True (sac.py:95-97)."""
    for k, v in fast.items():
        if k.split(".")[-1] in ("weight", "bias", "running_mean", "running_var"):
            slow[k].mul_(m).add_(v.detach() * (1.0 - m))


def sac_target_forward(student, teacher, running_conf, batch, T, cfg, training=True):
    """SAC.forward with use_teacher=True, teacher already initialised
    (sac.py:331-378). ``student``/``teacher``: state_dict-like mappings
    (student leaves may require grad). Returns (losses, outs, new running_conf)."""
    x, y, x2, affine, affine_inv = batch
    y = y.clone()
    ignore_mask = (y == -1)                                            # :337
    y[ignore_mask] = 255                                               # :338
    losses, outs = backbone_forward(student, x, y)                     # :340
    with torch.no_grad():
        slow_logits, _ = backbone_forward(teacher, x2)                 # :350
        refined, rc, diags = refine(slow_logits, x2.shape[-2:], T, affine, affine_inv,
                                    ignore_mask, running_conf, cfg, training)       # :353
        labels, conf, idx, thr = pseudo_labels_probs(refined, ignore_mask, rc, cfg, cfg.CONF_DISCOUNT)  # :357
    loss = focal_ce_conf(outs["logits_up"], labels, conf, rc, cfg.FOCAL_P)          # :360
    losses["self_ce"] = loss.view(1)                                                # :361
    with torch.no_grad():
        losses["teacher_diff"] = teacher_diff(teacher, student)                     # :374
    outs.update(teacher_logits=slow_logits, teacher_refined=refined, teacher_conf=conf,
                teacher_labels=labels, teacher_idx=idx, thresholds=thr, running_conf=rc,
                mask_gt=y, **diags)
    return losses, outs, rc


def parameter_groups(student, lr, wd):
    """BaseNet.parameter_groups for DeepLabV2_ResNet101 (basenet.py:102-139,
    deeplabv2.py:203-211): [old W (wd), old b, new W x10 (wd), new b x20]."""
    groups = ({"params": [], "weight_decay": wd, "lr": lr}, {"params": [], "weight_decay": 0.0, "lr": 2 * lr},
              {"params": [], "weight_decay": wd, "lr": 10 * lr}, {"params": [], "weight_decay": 0.0, "lr": 20 * lr})
    for k, v in student.items():
        if not (k.endswith(".weight") or k.endswith(".bias")):
            continue
        # from-scratch layers (deeplabv2.py:201, 285-287): ASPP head; for VGG also fc6 / fc7
        new = k.startswith(("model.layer5.", "classifier.", "features.42.", "features.44.", "vgg_head.", "score_pool"))
        isw = k.endswith(".weight")
        groups[(2 if new else 0) + (0 if isw else 1)]["params"].append(v)
    return list(groups)


def sac_target_step(student, teacher, running_conf, batch, T, cfg, optim=None):
    """Trainer._step_target with train=True, TARGET_ONLY semantics
    (train.py:211-233): forward, LR_TARGET*self_ce backward, SGD step."""
    losses, outs, rc = sac_target_forward(student, teacher, running_conf, batch, T, cfg, True)
    if optim is not None:
        optim.zero_grad()
    (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
    if optim is not None:
        optim.step()
    return losses, outs, rc


# --------------------------------------------------------------------------
# ABN baseline (cfg.MODEL.BASELINE = True): SAC_Baseline (models/sac.py:15-38) around a backbone whose BN layers train
# --------------------------------------------------------------------------

def baseline_forward(params, x, y=None, taps=None):
    """SAC_Baseline.forward -> backbone.forward in train() mode with freeze_bn=False (sac.py:34-35, deeplabv2.py:213-227,
    300-312, fcn.py:139-149; the architecture is recognised from the state_dict keys).  Returns (losses, outs, new_stats); ``new_stats`` holds the running statistics after the
    forward pass (every BN layer updates them, also under torch.no_grad())."""
    new_stats = OrderedDict()
    bn = bn_train_recorder(new_stats, taps=taps)
    if "model.conv1.weight" in params: logits = resnet101_logits(params, x, taps=taps, bn=bn)
    elif "vgg_head.0.weight" in params: logits = vgg16_fcn8s_logits(params, x, bn=bn)       # drop_rate = 0
    else: logits = vgg16_deeplab_logits(params, x, taps=taps, bn=bn)
    logits_up = F.interpolate(logits, x.shape[-2:], mode="bilinear", align_corners=True)
    if y is None:
        return (logits, logits_up), None, new_stats
    ce = F.cross_entropy(logits_up, y, ignore_index=255, reduction="none")
    return {"loss_ce": ce.mean().view(1)}, {"logits_up": logits_up, "logits": logits}, new_stats


def commit_bn_stats(params, new_stats):
    """the in-place buffer update of training-mode BN (+ num_batches_tracked, which the SAC path never reads)"""
    for k, v in new_stats.items():
        params[k] = v
    for k in list(params.keys()):
        if k.endswith(".num_batches_tracked") and (k[:-len("num_batches_tracked")] + "running_mean") in new_stats:
            params[k] = params[k] + 1


def baseline_source_step(params, x, y, optim=None):
    """Trainer.step(train=True) in BASELINE mode (train.py:119-138): forward with batch statistics, zero_grad,
    loss_ce.mean().backward(), optimiser step (immediately: train.py:134-138)."""
    losses, outs, new_stats = baseline_forward(params, x, y)
    if optim is not None:
        optim.zero_grad()
    losses["loss_ce"].mean().backward()
    if optim is not None:
        optim.step()
    commit_bn_stats(params, new_stats)
    return losses, outs


def baseline_target_pass(params, x, y):
    """The ABN target pass (train.py:113-115,281-289): ``step(train=False)`` under torch.no_grad() with the net in
    train() mode -- nothing is learnt, only the BN running statistics absorb the target batch."""
    with torch.no_grad():
        losses, outs, new_stats = baseline_forward(params, x, y)
    commit_bn_stats(params, new_stats)
    return losses, outs


def as_leaf_params(sd):
    """state_dict -> mapping whose weight/bias entries are autograd leaves."""
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith(".weight") or k.endswith(".bias"):
            out[k] = v.clone().requires_grad_(True)
        else:
            out[k] = v.clone()
    return out
