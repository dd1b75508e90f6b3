"""Layer executor for DeepLabV2-ResNet101 on libsac_b200 (host side, Python).

This is the part of the reference that lives in ``models/deeplabv2.py`` (ResNet /
Bottleneck / Classifier_Module forward) plus what autograd derives from it,
re-expressed as a fixed schedule of C-ABI kernel calls:

* activations: bf16 split planes, NHWC (include/sacb.h);
* every conv + frozen-BN (+ReLU) (+residual) unit is ONE ``sacb_conv_gemm`` launch;
* backward of a unit = one ``sacb_conv_gemm`` (data gradient, with the ReLU mask of the
  producer fused) + one ``sacb_conv_wgrad`` + ``sacb_colsum`` (d beta) +
  ``sacb_wgrad_finalize`` (dW re-layout, d gamma from <W, dW_raw>).

No PyTorch op touches the activations; torch provides memory and streams only.
"""
import ctypes as C
import os

import torch

from . import lib as L

# SACB_BWD_TWO_STREAM (default on, 0 = off): the filter-gradient GEMMs of the backward pass run on a side stream.  They are off the critical path
# (nothing but the final wgrad_finalize consumes them) while the data-gradient GEMMs form a dependent chain; every GEMM is a
# persistent kernel whose last partial wave and prologue / drain leave SMs idle, which the other stream's kernel fills.
BWD_TWO_STREAM = os.environ.get("SACB_BWD_TWO_STREAM", "1") != "0"

BN_EPS = 1e-5
NUM_CLASSES = 19
ASPP_JPAD = 768          # 4 convs x 9 taps x 19 classes = 684, padded to the GEMM N tile (sacb_aspp_jpad)
ASPP_DIL = (C.c_int32 * 4)(6, 12, 18, 24)
STEM_KP = 192            # 3*7*7 = 147 im2col columns of the stem, zero-padded to a multiple of 64


class ConvSpec(object):
    __slots__ = ("name", "bn", "K", "C", "R", "stride", "dil", "pad", "hin", "win", "hout", "wout", "Kf", "Kt", "bias", "head")

    def __init__(self, name, bn, K, Cc, R, stride, dil, pad, bias=False, head=False):
        self.name, self.bn, self.K, self.C, self.R = name, bn, K, Cc, R
        self.stride, self.dil, self.pad = stride, dil, pad
        self.bias = bias            # conv has its own bias parameter (VGG convs, ASPP head)
        self.head = head            # one of the four ASPP classifier convs (handled by the packed tap-unrolled GEMM)
        self.Kf = (K + 31) // 32 * 32       # rows of the fprop weight planes
        self.Kt = (K + 63) // 64 * 64       # reduction length of the dgrad weight planes

    def geom(self, N):
        return (N, self.hin, self.win, self.C, self.Kf, self.R, self.stride, self.dil, self.pad)

    def geom_dgrad(self, N):
        # data gradient of a stride-1 conv = conv of G with flipped/transposed weights, same pad/dilation
        return (N, self.hout, self.wout, self.Kt, self.C, self.R, 1, self.dil, self.pad if self.stride == 1 else 0)


def build_resnet101(H, W):
    """Layer table of ResNet(Bottleneck, [3,4,23,3]) + ASPP (/root/reference/models/deeplabv2.py:118-171)
    with spatial sizes resolved for an H x W input."""
    specs = {}
    order = []

    def add(name, bn, K, Cc, R, stride, dil, pad, hin, win):
        s = ConvSpec(name, bn, K, Cc, R, stride, dil, pad)
        s.hin, s.win = hin, win
        s.hout, s.wout = L.conv_out_hw(hin, win, R, stride, dil, pad)
        specs[name] = s
        order.append(name)
        return s

    stem = add("model.conv1", "model.bn1", 64, 3, 7, 2, 1, 3, H, W)
    ph = (stem.hout + 2 - 3 + 1) // 2 + 1
    pw = (stem.wout + 2 - 3 + 1) // 2 + 1
    if (ph - 1) * 2 >= stem.hout + 1: ph -= 1
    if (pw - 1) * 2 >= stem.wout + 1: pw -= 1
    blocks = []
    inplanes, h, w = 64, ph, pw
    for li, (planes, nblocks, stride, dil) in enumerate(((64, 3, 1, 1), (128, 4, 2, 1), (256, 23, 1, 2), (512, 3, 1, 4)), 1):
        for b in range(nblocks):
            p = "model.layer%d.%d" % (li, b)
            s = stride if b == 0 else 1
            c1 = add(p + ".conv1", p + ".bn1", planes, inplanes, 1, s, 1, 0, h, w)
            c2 = add(p + ".conv2", p + ".bn2", planes, planes, 3, 1, dil, dil, c1.hout, c1.wout)
            c3 = add(p + ".conv3", p + ".bn3", planes * 4, planes, 1, 1, 1, 0, c2.hout, c2.wout)
            ds = add(p + ".downsample.0", p + ".downsample.1", planes * 4, inplanes, 1, s, 1, 0, h, w) if b == 0 else None
            blocks.append((p, c1, c2, c3, ds))
            inplanes, h, w = planes * 4, c3.hout, c3.wout
    aspp = [add("model.layer5.conv2d_list.%d" % i, None, NUM_CLASSES, 2048, 3, 1, d, d, h, w)
            for i, d in enumerate((6, 12, 18, 24))]
    for a_ in aspp:
        a_.bias = True; a_.head = True
    return dict(arch="resnet101", specs=specs, order=order, stem=stem, pool_hw=(ph, pw), blocks=blocks, aspp=aspp, out_hw=(h, w),
                stem_kp=STEM_KP)


VGG16_CONVS = ((0, 3, 64, 1), (3, 64, 64, 1), (7, 64, 128, 1), (10, 128, 128, 1), (14, 128, 256, 1), (17, 256, 256, 1),
               (20, 256, 256, 1), (24, 256, 512, 1), (27, 512, 512, 1), (30, 512, 512, 1), (33, 512, 512, 2),
               (36, 512, 512, 2), (39, 512, 512, 2))
VGG16_POOL_AFTER = (3, 10, 20)          # features.6 / .13 / .23 follow these convs; pool4 / pool5 are removed


def build_vgg16_deeplab(H, W):
    """Layer table of DeepLabV2_VGG16(use_bn=True) (/root/reference/models/deeplabv2.py:229-312): torchvision
    vgg16_bn features with conv5 dilated by 2, pool4/pool5 removed, fc6/fc7 as 3x3 dilation-4 convs, ASPP on 1024 ch."""
    specs, order, seq = {}, [], []
    h, w = H, W
    for (idx, cin, cout, dil) in VGG16_CONVS:
        s = ConvSpec("features.%d" % idx, "features.%d" % (idx + 1), cout, cin, 3, 1, dil, dil, bias=True)
        s.hin, s.win, s.hout, s.wout = h, w, h, w
        specs[s.name] = s; order.append(s.name); seq.append(("conv", s))
        if idx in VGG16_POOL_AFTER:
            seq.append(("pool", (h, w, h // 2, w // 2, cout, "pool%d" % idx)))
            h, w = h // 2, w // 2
    for (idx, cin, cout) in ((42, 512, 1024), (44, 1024, 1024)):
        s = ConvSpec("features.%d" % idx, None, cout, cin, 3, 1, 4, 4, bias=True)
        s.hin, s.win, s.hout, s.wout = h, w, h, w
        specs[s.name] = s; order.append(s.name); seq.append(("conv", s))
    aspp = []
    for i, d in enumerate((6, 12, 18, 24)):
        s = ConvSpec("classifier.conv2d_list.%d" % i, None, NUM_CLASSES, 1024, 3, 1, d, d, bias=True, head=True)
        s.hin, s.win, s.hout, s.wout = h, w, h, w
        specs[s.name] = s; order.append(s.name); aspp.append(s)
    return dict(arch="vgg16", specs=specs, order=order, stem=seq[0][1], seq=seq, aspp=aspp, out_hw=(h, w), stem_kp=64)


FCN_BLOCKS = (("block1", ((0, 3, 64), (3, 64, 64), (7, 64, 128), (10, 128, 128), (14, 128, 256), (17, 256, 256), (20, 256, 256)), (3, 10, 20)),
              ("block2", ((24, 256, 512), (27, 512, 512), (30, 512, 512)), (30,)),
              ("block3", ((34, 512, 512), (37, 512, 512), (40, 512, 512)), (40,)))


def build_vgg16_fcn(H, W):
    """Layer table of VGG16_FCN8s(use_bn=True) (/root/reference/models/fcn.py:10-149): vgg16_bn features split into
    block1 (.. pool3), block2 (.. pool4), block3 (.. pool5); head = 7x7 conv 512->4096 + BN + ReLU + Dropout2d, 1x1
    4096->4096 + BN + ReLU + Dropout2d, 1x1 4096->19; score_pool4 / score_pool3 1x1 convs; bilinear x2 fusion."""
    specs, order, seq = {}, [], []
    h, w = H, W

    def add(name, bn, K, Cc, R, dil, pad):
        s = ConvSpec(name, bn, K, Cc, R, 1, dil, pad, bias=True)
        s.hin, s.win, s.hout, s.wout = h, w, h, w
        specs[name] = s; order.append(name)
        return s

    taps = {}
    for blk, convs, pools in FCN_BLOCKS:
        for (idx, cin, cout) in convs:
            s = add("%s.%d" % (blk, idx), "%s.%d" % (blk, idx + 1), cout, cin, 3, 1, 1)
            seq.append(("conv", s))
            if idx in pools:
                tag = "pool%d" % idx
                seq.append(("pool", (h, w, h // 2, w // 2, cout, tag)))
                h, w = h // 2, w // 2
                taps[tag] = (h, w, cout)
    head0 = add("vgg_head.0", "vgg_head.1", 4096, 512, 7, 1, 3)
    head4 = add("vgg_head.4", "vgg_head.5", 4096, 4096, 1, 1, 0)
    head8 = add("vgg_head.8", None, NUM_CLASSES, 4096, 1, 1, 0)
    h16, w16, _ = taps["pool30"]
    h8, w8, _ = taps["pool20"]
    sp4 = ConvSpec("score_pool4", None, NUM_CLASSES, 512, 1, 1, 1, 0, bias=True)
    sp4.hin, sp4.win, sp4.hout, sp4.wout = h16, w16, h16, w16
    sp3 = ConvSpec("score_pool3", None, NUM_CLASSES, 256, 1, 1, 1, 0, bias=True)
    sp3.hin, sp3.win, sp3.hout, sp3.wout = h8, w8, h8, w8
    for s_ in (sp4, sp3):
        specs[s_.name] = s_; order.append(s_.name)
    return dict(arch="fcn", specs=specs, order=order, stem=seq[0][1], seq=seq, aspp=[], out_hw=(h8, w8), stem_kp=64,
                head=(head0, head4, head8), score=(sp4, sp3), hw32=(h, w), hw16=(h16, w16))


def build_net(arch, H, W):
    return {"resnet101": build_resnet101, "vgg16": build_vgg16_deeplab, "fcn": build_vgg16_fcn}[arch](H, W)


def param_layout(net):
    """Flat-buffer layout in the reference's state_dict order: every float tensor whose key ends in
    weight / bias / running_mean / running_var (the set SAC._momentum_update walks, sac.py:88-91)."""
    entries = []   # (key, shape, is_param)
    for name in net["order"]:
        s = net["specs"][name]
        entries.append((name + ".weight", (s.K, s.C, s.R, s.R), True))
        if s.bias:
            entries.append((name + ".bias", (s.K,), True))
        if s.bn is not None:
            entries.append((s.bn + ".weight", (s.K,), True))
            entries.append((s.bn + ".bias", (s.K,), True))
            entries.append((s.bn + ".running_mean", (s.K,), False))
            entries.append((s.bn + ".running_var", (s.K,), False))
    offs, o = {}, 0
    for key, shape, is_p in entries:
        n = 1
        for d in shape: n *= d
        offs[key] = (o, n, shape, is_p)
        o += (n + 3) // 4 * 4       # keep every tensor 16-byte aligned
    return entries, offs, o


class FlatParams(object):
    """One contiguous fp32 buffer holding a backbone's parameters and BN statistics."""

    def __init__(self, net, device):
        self.entries, self.offs, self.total = param_layout(net)
        self.buf = torch.zeros(self.total, device=device, dtype=torch.float32)

    def view(self, key):
        o, n, shape, _ = self.offs[key]
        return self.buf[o:o + n].view(shape)

    def load(self, sd, prefix=""):
        for key, _, _ in self.entries:
            self.view(key).copy_(sd[prefix + key])

    def ranges(self, params_only=False):
        r = []
        for key, _, is_p in self.entries:
            if params_only and not is_p: continue
            o, n, _, _ = self.offs[key]
            r += [o, o + n]
        return r


class WeightPlanes(object):
    """bf16 split planes of every conv's weights in the layouts the GEMM kernels consume, plus the folded
    BN affine (scale/shift).  Rebuilt from a FlatParams by ``prepare`` (once per optimiser step for the
    student, once per EMA update for the teacher)."""

    def __init__(self, net, device, with_dgrad, fold_bn=True):
        self.net, self.with_dgrad = net, with_dgrad
        # fold_bn=False (ABN baseline, training-mode BN): the dgrad planes stay un-scaled and scale/shift are filled by the
        # batch-statistics kernels of engine_abn instead of the running statistics
        self.fold_bn = fold_bn
        nf = nt = nsc = 0
        self.off = {}
        for name in net["order"]:
            s = net["specs"][name]
            if s.C == 3 or s.head:             # first conv: packed im2col planes; ASPP head: packed tap-unrolled planes
                self.off[name] = (None, None, nsc)
            else:
                self.off[name] = (nf, nt, nsc)
                nf += s.R * s.R * s.Kf * s.C
                nt += s.R * s.R * s.C * s.Kt
            nsc += s.Kf
        self.stem_off = nf
        self.stem_n = net["stem"].K * net["stem_kp"]
        nf += self.stem_n
        self.aspp_off = (nf, nt)
        self.aspp_n = ASPP_JPAD * net["aspp"][0].C if net["aspp"] else 0
        nf += self.aspp_n; nt += self.aspp_n
        bf = torch.bfloat16
        self.wf_hi = torch.empty(nf, device=device, dtype=bf); self.wf_lo = torch.empty(nf, device=device, dtype=bf)
        if with_dgrad:
            self.wt_hi = torch.empty(nt, device=device, dtype=bf); self.wt_lo = torch.empty(nt, device=device, dtype=bf)
        self.scale = torch.ones(nsc, device=device); self.shift = torch.zeros(nsc, device=device)
        # units whose fprop planes carry the folded BN scale (SacbPrepItem.fold_wf): their epilogue runs with a unit scale, which
        # lets the pair kernel add a residual through the tensor core (SacbConvGemm.unit_scale); the true scale vector is still
        # produced -- wgrad_finalize needs gamma / sigma
        self.ones = torch.ones(max(s.Kf for s in net["specs"].values()), device=device)
        self.folded = set()

    def wf(self, name):
        s = self.net["specs"][name]; o = self.off[name][0]; n = s.R * s.R * s.Kf * s.C
        return self.wf_hi[o:o + n], self.wf_lo[o:o + n]

    def wt(self, name):
        s = self.net["specs"][name]; o = self.off[name][1]; n = s.R * s.R * s.C * s.Kt
        return self.wt_hi[o:o + n], self.wt_lo[o:o + n]

    def stem(self):
        o = self.stem_off
        return self.wf_hi[o:o + self.stem_n], self.wf_lo[o:o + self.stem_n]

    def aspp(self):
        of, ot = self.aspp_off
        n = self.aspp_n
        f = (self.wf_hi[of:of + n], self.wf_lo[of:of + n])
        t = (self.wt_hi[ot:ot + n], self.wt_lo[ot:ot + n]) if self.with_dgrad else (None, None)
        return f, t

    def affine(self, name):
        s = self.net["specs"][name]; o = self.off[name][2]
        return self.scale[o:o + s.Kf], self.shift[o:o + s.Kf]

    def epi_affine(self, name):
        """(scale, shift, unit_scale) for the fprop epilogue of a conv unit"""
        sc, sh = self.affine(name)
        if name in self.folded:
            return self.ones[:sc.numel()], sh, True
        return sc, sh, False

    def _prep_table(self, flat):
        """device descriptor table for sacb_prepare_batched (rebuilt if the flat buffer moved, e.g. into peer memory)"""
        key = flat.buf.data_ptr()
        if getattr(self, "_ptab", None) is not None and self._ptab[0] == key:
            return self._ptab[1]
        lib = L.lib()
        items, blocks = [], []
        for name in self.net["order"]:
            s = self.net["specs"][name]
            if s.head:
                continue
            sc, sh = self.affine(name)
            bn = s.bn is not None and self.fold_bn
            planes = s.C != 3
            fold_wf = 1 if (bn and planes) else 0
            if fold_wf:
                self.folded.add(name)
            fh, fl = self.wf(name) if planes else (None, None)
            th, tl = self.wt(name) if (planes and self.with_dgrad) else (None, None)
            items.append(L.PrepItem(L.dptr(flat.view(name + ".weight")),
                                    L.dptr(flat.view(s.bn + ".weight")) if bn else None,
                                    L.dptr(flat.view(s.bn + ".bias")) if bn else None,
                                    L.dptr(flat.view(s.bn + ".running_mean")) if bn else None,
                                    L.dptr(flat.view(s.bn + ".running_var")) if bn else None,
                                    L.dptr(flat.view(name + ".bias")) if s.bias else None,
                                    L.dptr(sc), L.dptr(sh), L.dptr(fh), L.dptr(fl), L.dptr(th), L.dptr(tl),
                                    s.K, s.C, s.R, s.R, s.Kf, s.Kt, fold_wf, 0))
            blocks.append(lib.sacb_prep_item_blocks(s.K, s.C, s.R, s.R, s.Kf, s.Kt, 1 if planes else 0,
                                                    1 if (planes and self.with_dgrad) else 0))
        self._ptab = (key, L.item_table(items, blocks, flat.buf.device))
        return self._ptab[1]

    def prepare(self, flat):
        lib, st = L.lib(), L.stream()
        # BN fold + fprop / dgrad weight planes of every conv unit: ONE launch (was 2 per layer)
        items, begin, n, total = self._prep_table(flat)
        L.check(lib.sacb_prepare_batched(L.ptr(items), L.ptr(begin), n, total, C.c_float(BN_EPS), st), "sacb_prepare_batched")
        sh_, sl_ = self.stem()
        stem = self.net["stem"]
        L.check(lib.sacb_stem_pack_weight(L.ptr(flat.view(stem.name + ".weight")), L.ptr(sh_), L.ptr(sl_), stem.K,
                                          3 * stem.R * stem.R, self.net["stem_kp"], st), "sacb_stem_pack_weight")
        a = self.net["aspp"]
        if not a:
            return
        f, t = self.aspp()
        L.check(lib.sacb_aspp_pack_weights(*[L.ptr(flat.view(x.name + ".weight")) for x in a], a[0].C,
                                           L.ptr(f[0]), L.ptr(f[1]), L.ptr(t[0]), L.ptr(t[1]), st), "sacb_aspp_pack_weights")


class Planes(object):
    """a bf16 hi/lo pair viewed as [M, C]; ``slot`` = index of the scratch-pool buffers it lives in (None: persistent planes)"""
    __slots__ = ("hi", "lo", "slot")

    def __init__(self, hi, lo, slot=None):
        self.hi, self.lo, self.slot = hi, lo, slot


class BufferPool(object):
    """fixed set of max-size scratch buffers handed out by name.  ``fifo``: the buffer that has been free longest is handed out
    next (two-stream backward: a freed buffer may still be read by a kernel on the side stream; ``readers[i]`` holds the events
    the next user has to wait for)"""

    def __init__(self, n_elems, count, dtype, device, fifo=False):
        self.bufs = [torch.empty(n_elems, device=device, dtype=dtype) for _ in range(count)]
        self.free = list(range(count))
        self.used = {}
        self.fifo = fifo
        self.readers = {}

    def get(self, tag, n):
        assert tag not in self.used, tag
        i = self.free.pop(0) if self.fifo else self.free.pop()
        for ev in self.readers.pop(i, ()):
            torch.cuda.current_stream().wait_event(ev)
        self.used[tag] = i
        return self.bufs[i][:n]

    def put(self, tag):
        self.free.append(self.used.pop(tag))

    def reset(self):
        self.free = list(range(len(self.bufs))); self.used = {}; self.readers = {}


class EngineBase(object):
    """Buffers and building blocks shared by the backbone schedules: first conv through an explicit im2col GEMM,
    conv+BN(+ReLU)(+residual) units, filter-gradient + finalisation, the tap-unrolled ASPP head."""

    def __init__(self, net, N, H, W, device):
        self.net, self.N, self.H, self.W, self.device = net, N, H, W, device
        self.act = {}
        self.dwraw = torch.empty(1 << 20, device=device)
        stem = net["stem"]
        self.stem_a = self._planes(N * stem.hout * stem.wout, net["stem_kp"])     # im2col matrix of the first conv (kept for wgrad)
        self.stem_dw = torch.empty(stem.K * 3 * stem.R * stem.R, device=device)
        oh, ow = net["out_hw"]
        if net["aspp"]:
            self.zbuf = torch.empty(N * oh * ow * ASPP_JPAD, device=device)       # tap-unrolled ASPP partial outputs

    def _planes(self, m, c):
        bf = torch.bfloat16
        return Planes(torch.empty(m * c, device=self.device, dtype=bf), torch.empty(m * c, device=self.device, dtype=bf))

    def _make_pools(self, max_elems, n_planes, n_f32):
        bf = torch.bfloat16
        self.max_elems = max_elems
        self._side = None
        if BWD_TWO_STREAM and L.on_device(self.dwraw):
            self._side = torch.cuda.Stream()
            n_planes += 4              # distance between freeing a gradient buffer and overwriting it: the side stream's slack
        fifo = self._side is not None
        self.tpool_hi = BufferPool(max_elems, n_planes, bf, self.device, fifo); self.tpool_lo = BufferPool(max_elems, n_planes, bf, self.device, fifo)
        self.fpool = BufferPool(max_elems, n_f32, torch.float32, self.device)

    def _tplanes(self, tag, n):
        hi = self.tpool_hi.get(tag, n)
        return Planes(hi, self.tpool_lo.get(tag, n), slot=(self.tpool_hi.used[tag], self.tpool_lo.used[tag]))

    def _tput(self, tag):
        self.tpool_hi.put(tag); self.tpool_lo.put(tag)

    def _rename(self, old, new):
        self.tpool_hi.used[new] = self.tpool_hi.used.pop(old); self.tpool_lo.used[new] = self.tpool_lo.used.pop(old)

    # ---- forward pieces
    def _first_conv_fwd(self, flat, wp, x, out):
        lib, st, N, stem = L.lib(), L.stream(), self.N, self.net["stem"]
        kp = self.net["stem_kp"]
        L.check(lib.sacb_stem_im2col(L.ptr(x), L.ptr(self.stem_a.hi), L.ptr(self.stem_a.lo), N, self.H, self.W,
                                     stem.hout, stem.wout, stem.R, stem.stride, stem.pad, kp, st), "sacb_stem_im2col")
        sc, sh = wp.affine(stem.name)
        wsh, wsl = wp.stem()
        L.conv_gemm(self.stem_a.hi, self.stem_a.lo, wsh, wsl, (N, stem.hout, stem.wout, kp, stem.K, 1, 1, 1, 0),
                    scale=sc, shift=sh, relu=True, out_hi=out.hi, out_lo=out.lo)

    def _unit(self, wp, s, xin, out, relu, res=None):
        fh, fl = wp.wf(s.name); sc, sh, unit = wp.epi_affine(s.name)
        L.conv_gemm(xin.hi, xin.lo, fh, fl, s.geom(self.N), scale=sc, shift=sh,
                    add_hi=None if res is None else res.hi, add_lo=None if res is None else res.lo,
                    relu=relu, out_hi=out.hi, out_lo=out.lo, unit_scale=unit)

    def _aspp_fwd(self, flat, wp, a, logits_out):
        """ASPP head (deeplabv2.py:112-116) as one tap-unrolled 1x1 GEMM + shift-and-add (csrc/sacb_aspp.cu)"""
        lib, st, N, asp = L.lib(), L.stream(), self.N, self.net["aspp"]
        oh, ow = self.net["out_hw"]
        (fh, fl), _ = wp.aspp()
        L.conv_gemm(a.hi, a.lo, fh, fl, (N, oh, ow, asp[0].C, ASPP_JPAD, 1, 1, 1, 0), out_f32=self.zbuf)
        L.check(lib.sacb_aspp_gather(L.ptr(self.zbuf), *[L.ptr(flat.view(x.name + ".bias")) for x in asp], ASPP_DIL,
                                     L.ptr(logits_out), N, oh, ow, st), "sacb_aspp_gather")

    # ---- backward pieces
    def _begin_backward(self):
        self.tpool_hi.reset(); self.tpool_lo.reset(); self.fpool.reset()
        # d(beta) of every BN unit (d(bias) of bias-only convs) = column sums of the gradient arriving at it; accumulated by
        # the epilogue of the GEMM that produces that gradient (colsum=...), into one zero-filled buffer
        if getattr(self, "_dbeta_pool", None) is None:      # persistent: the batched finalize table holds pointers into it
            self._dbeta_pool = torch.zeros(sum(s.K for s in self.net["specs"].values() if not s.head) + 4096, device=self.device)
        else:
            self._dbeta_pool.zero_()
        self._dbeta_off = 0
        self._fin_pending = []

    def _aspp_bwd(self, flat, wp, xlast, dlogits, grad):
        """Gcol (shifted copies of dlogits) -> bias / filter / data gradients as plain GEMMs. Returns (gout, dbeta of xlast's unit)."""
        lib, st, N, asp = L.lib(), L.stream(), self.N, self.net["aspp"]
        oh, ow = self.net["out_hw"]
        M5 = N * oh * ow
        gcol = self._tplanes("gcol", M5 * ASPP_JPAD)
        L.check(lib.sacb_aspp_gcol(L.ptr(dlogits), ASPP_DIL, L.ptr(gcol.hi), L.ptr(gcol.lo), N, oh, ow, st), "sacb_aspp_gcol")
        csum = self._dbeta(gcol, M5, ASPP_JPAD)
        for i, s in enumerate(asp):                      # centre tap of conv i carries g itself: its column sum is d bias
            o = (i * 9 + 4) * NUM_CLASSES
            grad.view(s.name + ".bias").copy_(csum[o:o + NUM_CLASSES])
        parts, splits = L.conv_wgrad(xlast.hi, xlast.lo, gcol.hi, gcol.lo, self._dw_workspace,
                                     (N, oh, ow, asp[0].C, ASPP_JPAD, 1, 1, 1, 0))
        L.check(lib.sacb_aspp_unpack_wgrad(L.ptr(parts), splits, asp[0].C, *[L.ptr(grad.view(x.name + ".weight")) for x in asp], st),
                "sacb_aspp_unpack_wgrad")
        gout = self._tplanes("gout", M5 * asp[0].C)
        _, (th, tl) = wp.aspp()
        dbeta = self._new_dbeta(asp[0].C)
        L.conv_gemm(gcol.hi, gcol.lo, th, tl, (N, oh, ow, ASPP_JPAD, asp[0].C, 1, 1, 1, 0), mask_hi=xlast.hi,
                    out_hi=gout.hi, out_lo=gout.lo, colsum=dbeta)
        self._tput("gcol")
        return gout, dbeta

    def _first_conv_bwd(self, flat, wp, gs, grad, dbeta):
        lib, st, N, stem = L.lib(), L.stream(), self.N, self.net["stem"]
        kp, taps = self.net["stem_kp"], 3 * stem.R * stem.R
        parts, splits = L.conv_wgrad(self.stem_a.hi, self.stem_a.lo, gs.hi, gs.lo, self._dw_workspace,
                                     (N, stem.hout, stem.wout, kp, stem.Kt, 1, 1, 1, 0), k_valid=stem.K)
        L.check(lib.sacb_stem_unpack_wgrad(L.ptr(parts), splits, L.ptr(self.stem_dw), stem.K, taps, kp, st), "sacb_stem_unpack_wgrad")
        self._finalize(flat, wp, stem, self.stem_dw, grad, dbeta, C_eff=taps, RS=1)

    def _trunk_backward(self, flat, wp, seq, i, g, dbeta, gp, grad, inject=None):
        """Backward through a plain conv / 2x2-max-pool chain seq[0..i] (VGG trunks).
        Entering at a conv: ``g`` = masked gradient planes at its output (pool tag "gout"), ``dbeta`` their column sums.
        Entering at a pool: ``gp`` = un-masked fp32 gradient at the pool output. ``inject[tag]`` (fp32) is added to the
        gradient arriving at pool ``tag`` (FCN score branches)."""
        lib, st, N = L.lib(), L.stream(), self.N
        inject = inject or {}
        while i >= 0:
            kind, it = seq[i]
            if kind == "pool":
                h, w, ph, pw, c, tag = it
                src = seq[i - 1][1]                       # conv feeding the pool (its output is a ReLU output)
                gx = self._tplanes("gout", N * h * w * c)
                L.check(lib.sacb_maxpool2_bwd(L.ptr(gp), L.ptr(self.pool_idx[tag]), L.ptr(self.act[src.name].hi), L.ptr(gx.hi),
                                              L.ptr(gx.lo), N, h, w, c, ph, pw, st), "sacb_maxpool2_bwd")
                self.fpool.put("gp")
                dbeta = self._new_dbeta(c)
                L.check(lib.sacb_colsum(L.ptr(gx.hi), L.ptr(gx.lo), L.ptr(dbeta), C.c_int64(N * h * w), c, st), "sacb_colsum")
                g = gx
                i -= 1
                continue
            s = it
            if i == 0:
                self._first_conv_bwd(flat, wp, g, grad, dbeta)
                self._tput("gout")
                return
            pkind, pit = seq[i - 1]
            xin = self.act[pit.name] if pkind == "conv" else self.act[pit[5]]
            self._wgrad(flat, wp, s, xin, g, grad, dbeta)
            th, tl = wp.wt(s.name)
            if pkind == "conv":
                gx = self._tplanes("gx", N * s.hin * s.win * s.C)
                dbeta = self._new_dbeta(s.C)
                L.conv_gemm(g.hi, g.lo, th, tl, s.geom_dgrad(N), mask_hi=xin.hi, out_hi=gx.hi, out_lo=gx.lo, colsum=dbeta)
                self._tput("gout"); self._rename("gx", "gout")
                g = gx
            else:
                # the conv's input is a max-pool output: un-masked fp32 gradient, routed through the pool afterwards
                gp = self.fpool.get("gp", N * s.hin * s.win * s.C)
                L.conv_gemm(g.hi, g.lo, th, tl, s.geom_dgrad(N), add_f32=inject.get(pit[5]), out_f32=gp)
                self._tput("gout")
            i -= 1

    def _new_dbeta(self, K):
        o = self._dbeta_off
        self._dbeta_off = o + K
        assert self._dbeta_off <= self._dbeta_pool.numel(), "d(beta) pool exhausted"
        return self._dbeta_pool[o:o + K]

    def _dbeta(self, g, M, K):
        d = self._new_dbeta(K)                 # slice of the zero-filled persistent pool
        L.check(L.lib().sacb_colsum(L.ptr(g.hi), L.ptr(g.lo), L.ptr(d), C.c_int64(M), K, L.stream()), "sacb_colsum")
        return d

    def _dw_workspace(self, n):
        if self.dwraw.numel() < n:
            self.dwraw = torch.empty(n, device=self.device)
        return self.dwraw

    def _dw_for(self, name, n):
        """split-K partial planes of one layer: every layer keeps its own region until the batched finalize at the end
        of the backward pass (~38 MB per layer: one 256x256 fp32 tile per CTA pair)"""
        arena = self.__dict__.setdefault("_dw_arena", {})
        t = arena.get(name)
        if t is None or t.numel() < n:
            t = arena[name] = torch.empty(n, device=self.device)
            self._fin_table = None
        return t

    def _wgrad(self, flat, wp, s, xin, g, grad, dbeta):
        geom = (self.N, s.hin, s.win, s.C, s.Kt, s.R, s.stride, s.dil, s.pad)
        side = None if L.SERIALIZE else getattr(self, "_side", None)
        if side is None:
            dwraw, splits = L.conv_wgrad(xin.hi, xin.lo, g.hi, g.lo, lambda n: self._dw_for(s.name, n), geom, k_valid=s.K)
        else:
            # side stream: ordered after everything issued so far (g and xin are complete), its result is only read by the
            # batched finalize after the join; whoever re-uses g's (or xin's) scratch buffers first waits for this kernel
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                dwraw, splits = L.conv_wgrad(xin.hi, xin.lo, g.hi, g.lo, lambda n: self._dw_for(s.name, n), geom, k_valid=s.K)
                done = torch.cuda.Event()
                done.record(side)
            for pl in (g, xin):
                if pl.slot is not None:
                    self.tpool_hi.readers.setdefault(pl.slot[0], []).append(done)
                    self.tpool_lo.readers.setdefault(pl.slot[1], []).append(done)
        self._finalize(flat, wp, s, dwraw, grad, dbeta, C_eff=s.C, RS=s.R * s.R, splits=splits)

    def _finalize(self, flat, wp, s, dwraw, grad, dbeta, C_eff, RS, splits=1):
        """dW (OIHW), d gamma, d beta and the conv-bias gradient of one unit from its raw filter gradient and d beta.
        Deferred: all units of a backward pass are finalised by ONE launch (``_finalize_all``)."""
        self._fin_pending.append((s, dwraw, dbeta, C_eff, RS, splits))

    def _finalize_all(self, flat, wp, grad):
        if getattr(self, "_side", None) is not None:
            torch.cuda.current_stream().wait_stream(self._side)      # join: every filter gradient has been written
        key = (flat.buf.data_ptr(), grad.buf.data_ptr(), wp.scale.data_ptr(), len(self._fin_pending))
        tab = getattr(self, "_fin_table", None)
        if tab is None or tab[0] != key:
            items, blocks = [], []
            for s, dwraw, dbeta, C_eff, RS, splits in self._fin_pending:
                bn = s.bn is not None
                sc = wp.affine(s.name)[0] if bn else None
                items.append(L.FinalizeItem(
                    L.dptr(dwraw), L.dptr(flat.view(s.name + ".weight")), L.dptr(sc),
                    L.dptr(flat.view(s.bn + ".running_mean")) if bn else None,
                    L.dptr(flat.view(s.bn + ".running_var")) if bn else None, L.dptr(dbeta),
                    L.dptr(grad.view(s.name + ".weight")), L.dptr(grad.view(s.bn + ".weight")) if bn else None,
                    L.dptr(flat.view(s.name + ".bias")) if s.bias else None,
                    L.dptr(grad.view(s.name + ".bias")) if s.bias else None,
                    L.dptr(grad.view(s.bn + ".bias")) if bn else None, s.K, C_eff, RS, splits))
                blocks.append(s.K)
            tab = self._fin_table = (key, L.item_table(items, blocks, self.device))
        items, begin, n, total = tab[1]
        L.check(L.lib().sacb_wgrad_finalize_batched(L.ptr(items), L.ptr(begin), n, total, C.c_float(BN_EPS), L.stream()),
                "sacb_wgrad_finalize_batched")
        self._fin_pending = []


class ResNet101Engine(EngineBase):
    def __init__(self, N, H, W, device):
        EngineBase.__init__(self, build_resnet101(H, W), N, H, W, device)
        net = self.net
        st = net["stem"]
        ph, pw = net["pool_hw"]
        self.pool_idx = torch.empty(N * ph * pw * 64, device=device, dtype=torch.uint8)
        max_elems = N * st.hout * st.wout * 64
        for (p, c1, c2, c3, ds) in net["blocks"]:
            for c in (c1, c2, c3):
                max_elems = max(max_elems, N * c.hout * c.wout * c.K)
        self._make_pools(max_elems, 5, 2)

    def _ensure_act(self):
        """the persistent activation set of a forward pass that will be back-propagated (about 1 GB per 512^2 crop): allocated
        on the first such pass, so that an engine that only ever runs the no-grad teacher / inference forward (scratch planes)
        does not hold it"""
        if self.act:
            return
        net, N = self.net, self.N
        st = net["stem"]
        ph, pw = net["pool_hw"]
        self.act["stem"] = self._planes(N * st.hout * st.wout, 64)
        self.act["pool"] = self._planes(N * ph * pw, 64)
        for (p, c1, c2, c3, ds) in net["blocks"]:
            for c in (c1, c2, c3):
                self.act[c.name] = self._planes(N * c.hout * c.wout, c.K)
            if ds is not None:
                self.act[ds.name] = self._planes(N * ds.hout * ds.wout, ds.K)

    # ------------------------------------------------------------------ forward
    def forward(self, flat, wp, x, logits_out, keep):
        """x: fp32 NCHW [N,3,H,W]; logits_out: fp32 NCHW [N,19,h,w]. keep=True stores activations for backward."""
        net, N, lib, st = self.net, self.N, L.lib(), L.stream()
        stem = net["stem"]
        ph, pw = net["pool_hw"]
        self.tpool_hi.reset(); self.tpool_lo.reset()
        if keep:
            self._ensure_act()
        get = (lambda tag, n: self.act[tag]) if keep else self._tplanes
        put = (lambda tag: None) if keep else self._tput
        a_stem = get("stem", N * stem.hout * stem.wout * 64)
        self._first_conv_fwd(flat, wp, x, a_stem)
        a = get("pool", N * ph * pw * 64)
        L.check(lib.sacb_maxpool_fwd(L.ptr(a_stem.hi), L.ptr(a_stem.lo), L.ptr(a.hi), L.ptr(a.lo), L.ptr(self.pool_idx),
                                     N, stem.hout, stem.wout, 64, ph, pw, st), "sacb_maxpool_fwd")
        put("stem")
        xtag = "pool"
        for (p, c1, c2, c3, ds) in net["blocks"]:
            o1 = get(c1.name, N * c1.hout * c1.wout * c1.K)
            self._unit(wp, c1, a, o1, relu=True)
            o2 = get(c2.name, N * c2.hout * c2.wout * c2.K)
            self._unit(wp, c2, o1, o2, relu=True)
            put(c1.name)
            if ds is not None:
                r = get(ds.name, N * ds.hout * ds.wout * ds.K)
                self._unit(wp, ds, a, r, relu=False)
            else:
                r = a
            o3 = get(c3.name, N * c3.hout * c3.wout * c3.K)
            self._unit(wp, c3, o2, o3, relu=True, res=r)
            put(c2.name)
            if ds is not None: put(ds.name)
            put(xtag)
            a, xtag = o3, c3.name
        self._aspp_fwd(flat, wp, a, logits_out)
        put(xtag)
        return logits_out

    # ------------------------------------------------------------------ backward
    def backward(self, flat, wp, x, dlogits, grad):
        """dlogits fp32 NCHW [N,19,h,w]; writes every parameter gradient into ``grad`` (FlatParams layout)."""
        net, N, lib, st = self.net, self.N, L.lib(), L.stream()
        self._begin_backward()
        blocks = net["blocks"]
        gout, dbeta3 = self._aspp_bwd(flat, wp, self.act[blocks[-1][3].name], dlogits, grad)
        for bi in range(len(blocks) - 1, -1, -1):
            (p, c1, c2, c3, ds) = blocks[bi]
            xin = self.act[blocks[bi - 1][3].name] if bi > 0 else self.act["pool"]
            o1, o2 = self.act[c1.name], self.act[c2.name]
            M = N * c3.hout * c3.wout
            # conv3 + bn3 (no ReLU between bn3 and the residual sum): g3 = gout, its column sums are dbeta3
            self._wgrad(flat, wp, c3, o2, gout, grad, dbeta3)
            g2 = self._tplanes("g2", M * c2.K)
            dbeta2 = self._new_dbeta(c2.K)
            th, tl = wp.wt(c3.name)
            L.conv_gemm(gout.hi, gout.lo, th, tl, c3.geom_dgrad(N), mask_hi=o2.hi, out_hi=g2.hi, out_lo=g2.lo, colsum=dbeta2)
            # conv2 + bn2 + relu
            self._wgrad(flat, wp, c2, o1, g2, grad, dbeta2)
            g1 = self._tplanes("g1", M * c1.K)
            dbeta1 = self._new_dbeta(c1.K)
            th, tl = wp.wt(c2.name)
            L.conv_gemm(g2.hi, g2.lo, th, tl, c2.geom_dgrad(N), mask_hi=o1.hi, out_hi=g1.hi, out_lo=g1.lo, colsum=dbeta1)
            self._tput("g2")
            # conv1 + bn1 + relu (and the downsample branch of the first block of a layer)
            self._wgrad(flat, wp, c1, xin, g1, grad, dbeta1)
            if ds is not None:
                self._wgrad(flat, wp, ds, xin, gout, grad, dbeta3)      # d beta of the downsample BN == d beta of bn3
            Min = N * c1.hin * c1.win
            th1, tl1 = wp.wt(c1.name)
            if bi == 0:
                # block input is the max-pool output: no ReLU mask, gradient continues through the pool
                thd, tld = wp.wt(ds.name)
                tmp = self.fpool.get("tmp", Min * c1.C)
                L.conv_gemm(gout.hi, gout.lo, thd, tld, ds.geom_dgrad(N), out_f32=tmp)
                gp = self.fpool.get("gpool", Min * c1.C)
                L.conv_gemm(g1.hi, g1.lo, th1, tl1, c1.geom_dgrad(N), add_f32=tmp, out_f32=gp)
                self._tput("g1"); self._tput("gout")
                break
            gx = self._tplanes("gx", Min * c1.C)
            dbeta3 = self._new_dbeta(c1.C)          # d beta of the previous block's bn3 (and of its downsample BN)
            if ds is None:
                L.conv_gemm(g1.hi, g1.lo, th1, tl1, c1.geom_dgrad(N), add_hi=gout.hi, add_lo=gout.lo, mask_hi=xin.hi,
                            out_hi=gx.hi, out_lo=gx.lo, colsum=dbeta3)
            elif c1.stride == 1:
                thd, tld = wp.wt(ds.name)
                tmp = self.fpool.get("tmp", Min * c1.C)
                L.conv_gemm(gout.hi, gout.lo, thd, tld, ds.geom_dgrad(N), out_f32=tmp)
                L.conv_gemm(g1.hi, g1.lo, th1, tl1, c1.geom_dgrad(N), add_f32=tmp, mask_hi=xin.hi, out_hi=gx.hi, out_lo=gx.lo,
                            colsum=dbeta3)
                self.fpool.put("tmp")
            else:
                # stride-2 1x1 convs: compact data gradients on the 65x65 grid, scattered to the even pixels
                thd, tld = wp.wt(ds.name)
                ta = self.fpool.get("tmp", M * c1.C); tb = self.fpool.get("tmp2", M * c1.C)
                L.conv_gemm(g1.hi, g1.lo, th1, tl1, c1.geom_dgrad(N), out_f32=ta)
                L.conv_gemm(gout.hi, gout.lo, thd, tld, ds.geom_dgrad(N), out_f32=tb)
                L.check(lib.sacb_scatter2_mask_split(L.ptr(ta), L.ptr(tb), L.ptr(xin.hi), L.ptr(gx.hi), L.ptr(gx.lo),
                                                     N, c1.hin, c1.win, c1.C, c1.hout, c1.wout, st), "sacb_scatter2_mask_split")
                L.check(lib.sacb_colsum(L.ptr(gx.hi), L.ptr(gx.lo), L.ptr(dbeta3), C.c_int64(Min), c1.C, st), "sacb_colsum")
                self.fpool.put("tmp"); self.fpool.put("tmp2")
            self._tput("g1"); self._tput("gout")
            # rename gx -> gout for the next (earlier) block
            self._rename("gx", "gout")
            gout = gx
        # max-pool backward (+ ReLU mask of the stem) and the stem conv
        stem = net["stem"]
        ph, pw = net["pool_hw"]
        a_stem = self.act["stem"]
        gs = self._tplanes("gstem", N * stem.hout * stem.wout * 64)
        L.check(lib.sacb_maxpool_bwd(L.ptr(gp), L.ptr(self.pool_idx), L.ptr(a_stem.hi), L.ptr(gs.hi), L.ptr(gs.lo),
                                     N, stem.hout, stem.wout, 64, ph, pw, st), "sacb_maxpool_bwd")
        dbeta = self._dbeta(gs, N * stem.hout * stem.wout, 64)
        self._first_conv_bwd(flat, wp, gs, grad, dbeta)
        self._finalize_all(flat, wp, grad)


class VGG16Engine(EngineBase):
    """DeepLabV2_VGG16(use_bn=True): 13 conv+bias+BN+ReLU units with three 2x2 max-pools, fc6/fc7 (conv+bias+ReLU), ASPP."""

    def __init__(self, N, H, W, device):
        EngineBase.__init__(self, build_vgg16_deeplab(H, W), N, H, W, device)
        max_elems = 0
        self.pool_idx = {}
        for kind, it in self.net["seq"]:
            if kind == "conv":
                self.act[it.name] = self._planes(N * it.hout * it.wout, it.K)
                max_elems = max(max_elems, N * it.hout * it.wout * it.K)
            else:
                h, w, ph, pw, c, tag = it
                self.act[tag] = self._planes(N * ph * pw, c)
                self.pool_idx[tag] = torch.empty(N * ph * pw * c, device=device, dtype=torch.uint8)
        self._make_pools(max_elems, 4, 2)

    def forward(self, flat, wp, x, logits_out, keep):
        net, N, lib, st = self.net, self.N, L.lib(), L.stream()
        self.tpool_hi.reset(); self.tpool_lo.reset()
        get = (lambda tag, n: self.act[tag]) if keep else self._tplanes
        put = (lambda tag: None) if keep else self._tput
        a, atag = None, None
        for kind, it in net["seq"]:
            if kind == "conv":
                o = get(it.name, N * it.hout * it.wout * it.K)
                if a is None:
                    self._first_conv_fwd(flat, wp, x, o)
                else:
                    self._unit(wp, it, a, o, relu=True)
                    put(atag)
                a, atag = o, it.name
            else:
                h, w, ph, pw, c, tag = it
                o = get(tag, N * ph * pw * c)
                L.check(lib.sacb_maxpool2_fwd(L.ptr(a.hi), L.ptr(a.lo), L.ptr(o.hi), L.ptr(o.lo), L.ptr(self.pool_idx[tag]),
                                              N, h, w, c, ph, pw, st), "sacb_maxpool2_fwd")
                put(atag)
                a, atag = o, tag
        self._aspp_fwd(flat, wp, a, logits_out)
        put(atag)
        return logits_out

    def backward(self, flat, wp, x, dlogits, grad):
        net, N, lib, st = self.net, self.N, L.lib(), L.stream()
        self._begin_backward()
        seq = net["seq"]
        last = seq[-1][1]
        g, dbeta = self._aspp_bwd(flat, wp, self.act[last.name], dlogits, grad)      # g = grad at fc7's output (masked)
        self._trunk_backward(flat, wp, seq, len(seq) - 1, g, dbeta, None, grad)
        self._finalize_all(flat, wp, grad)


class FCN8sEngine(EngineBase):
    """VGG16_FCN8s(use_bn=True) (/root/reference/models/fcn.py): VGG trunk with five 2x2 pools, 7x7 / 1x1 head with
    channel dropout, three 19-class 1x1 score convs fused by bilinear x2 up-sampling (align_corners=True)."""

    def __init__(self, N, H, W, device):
        EngineBase.__init__(self, build_vgg16_fcn(H, W), N, H, W, device)
        net = self.net
        max_elems = 0
        self.pool_idx = {}
        for kind, it in net["seq"]:
            if kind == "conv":
                self.act[it.name] = self._planes(N * it.hout * it.wout, it.K)
                max_elems = max(max_elems, N * it.hout * it.wout * it.K)
            else:
                h, w, ph, pw, c, tag = it
                self.act[tag] = self._planes(N * ph * pw, c)
                self.pool_idx[tag] = torch.empty(N * ph * pw * c, device=device, dtype=torch.uint8)
        h32, w32 = net["hw32"]; h16, w16 = net["hw16"]; h8, w8 = net["out_hw"]
        head0, head4, head8 = net["head"]
        for hd in (head0, head4):
            self.act[hd.name] = self._planes(N * h32 * w32, hd.K)
        max_elems = max(max_elems, N * h32 * w32 * 4096)
        self._make_pools(max_elems, 4, 2)
        f32 = dict(device=device, dtype=torch.float32)
        self.s32 = torch.empty(N, NUM_CLASSES, h32, w32, **f32)
        self.s4 = torch.empty(N, NUM_CLASSES, h16, w16, **f32); self.t16 = torch.empty_like(self.s4)
        self.s3 = torch.empty(N, NUM_CLASSES, h8, w8, **f32)
        self.d16 = torch.empty_like(self.s4); self.d32 = torch.empty_like(self.s32)
        self.gsc = {8: self._planes(N * h8 * w8, 64), 16: self._planes(N * h16 * w16, 64), 32: self._planes(N * h32 * w32, 64)}
        self.inj3 = torch.empty(N * h8 * w8 * 256, **f32); self.inj4 = torch.empty(N * h16 * w16 * 512, **f32)
        self.dropout = None          # (mask0 [N,4096], mask4 [N,4096]) of Dropout2d scale factors, or None (eval / p = 0)

    def _score(self, wp, s, xin, out_nchw):
        fh, fl = wp.wf(s.name); sc, sh = wp.affine(s.name)
        L.conv_gemm(xin.hi, xin.lo, fh, fl, s.geom(self.N), k_valid=s.K, scale=sc, shift=sh, out_nchw=out_nchw)

    def forward(self, flat, wp, x, logits_out, keep):
        # teacher and student both use the persistent activation set: the (no-grad) teacher pass always precedes the
        # student pass of the same step, and the score convs need pool3 / pool4 after the trunk has moved on
        net, N, lib, st = self.net, self.N, L.lib(), L.stream()
        a = None
        for kind, it in net["seq"]:
            if kind == "conv":
                o = self.act[it.name]
                if a is None: self._first_conv_fwd(flat, wp, x, o)
                else: self._unit(wp, it, a, o, relu=True)
                a = o
            else:
                h, w, ph, pw, c, tag = it
                o = self.act[tag]
                L.check(lib.sacb_maxpool2_fwd(L.ptr(a.hi), L.ptr(a.lo), L.ptr(o.hi), L.ptr(o.lo), L.ptr(self.pool_idx[tag]),
                                              N, h, w, c, ph, pw, st), "sacb_maxpool2_fwd")
                a = o
        head0, head4, head8 = net["head"]
        sp4, sp3 = net["score"]
        h32, w32 = net["hw32"]; h16, w16 = net["hw16"]; h8, w8 = net["out_hw"]
        for j, hd in enumerate((head0, head4)):
            o = self.act[hd.name]
            self._unit(wp, hd, a, o, relu=True)
            if self.dropout is not None:                  # Dropout2d after the ReLU (fcn.py:52,56)
                L.check(lib.sacb_channel_scale(L.ptr(o.hi), L.ptr(o.lo), L.ptr(self.dropout[j]), N, h32, w32, hd.K, st), "sacb_channel_scale")
            a = o
        self._score(wp, head8, a, self.s32)
        self._score(wp, sp4, self.act["pool30"], self.s4)
        self._score(wp, sp3, self.act["pool20"], self.s3)
        Cn = NUM_CLASSES                                   # score fusion (fcn.py:111-134)
        L.check(lib.sacb_upsample_add(L.ptr(self.s32), L.ptr(self.s4), L.ptr(self.t16), N, Cn, h32, w32, h16, w16, st), "sacb_upsample_add")
        L.check(lib.sacb_upsample_add(L.ptr(self.t16), L.ptr(self.s3), L.ptr(logits_out), N, Cn, h16, w16, h8, w8, st), "sacb_upsample_add")
        return logits_out

    def _score_bwd(self, flat, wp, s, xin, d_nchw, res, grad, out_f32=None, mask=None, out=None, colsum=None):
        """19-class 1x1 conv: d bias, filter gradient, data gradient (fp32 for pool inputs, masked planes for the head)"""
        lib, st, N = L.lib(), L.stream(), self.N
        g = self.gsc[res]
        L.check(lib.sacb_nchw_to_planes(L.ptr(d_nchw), L.ptr(g.hi), L.ptr(g.lo), N, NUM_CLASSES, s.hout, s.wout, 64, st), "sacb_nchw_to_planes")
        db = self._dbeta(g, N * s.hout * s.wout, 64)
        self._wgrad(flat, wp, s, xin, g, grad, db)         # no BN: finalize copies scale(=1) * db into the bias gradient
        th, tl = wp.wt(s.name)
        L.conv_gemm(g.hi, g.lo, th, tl, s.geom_dgrad(N), out_f32=out_f32, mask_hi=mask,
                    out_hi=None if out is None else out.hi, out_lo=None if out is None else out.lo, colsum=colsum)

    def backward(self, flat, wp, x, dlogits, grad):
        net, N, lib, st = self.net, self.N, L.lib(), L.stream()
        self._begin_backward()
        head0, head4, head8 = net["head"]
        sp4, sp3 = net["score"]
        h32, w32 = net["hw32"]; h16, w16 = net["hw16"]; h8, w8 = net["out_hw"]
        Cn = NUM_CLASSES
        p3, p4, p5 = self.act["pool20"], self.act["pool30"], self.act["pool40"]
        a0, a4 = self.act[head0.name], self.act[head4.name]
        # score branches (adjoints of the two bilinear x2 fusions)
        self._score_bwd(flat, wp, sp3, p3, dlogits, 8, grad, out_f32=self.inj3)
        L.check(lib.sacb_upsample_bwd(L.ptr(dlogits), L.ptr(self.d16), N, Cn, h16, w16, h8, w8, st), "sacb_upsample_bwd")
        self._score_bwd(flat, wp, sp4, p4, self.d16, 16, grad, out_f32=self.inj4)
        L.check(lib.sacb_upsample_bwd(L.ptr(self.d16), L.ptr(self.d32), N, Cn, h32, w32, h16, w16, st), "sacb_upsample_bwd")
        # head: 4096->19, then the two conv+BN+ReLU(+dropout) units
        M32 = N * h32 * w32
        drop = self.dropout is not None
        g4 = self._tplanes("gout", M32 * head4.K)
        db4 = self._new_dbeta(head4.K)
        self._score_bwd(flat, wp, head8, a4, self.d32, 32, grad, mask=a4.hi, out=g4, colsum=None if drop else db4)
        if drop:
            L.check(lib.sacb_channel_scale(L.ptr(g4.hi), L.ptr(g4.lo), L.ptr(self.dropout[1]), N, h32, w32, head4.K, st), "sacb_channel_scale")
            L.check(lib.sacb_colsum(L.ptr(g4.hi), L.ptr(g4.lo), L.ptr(db4), C.c_int64(M32), head4.K, st), "sacb_colsum")
        self._wgrad(flat, wp, head4, a0, g4, grad, db4)
        g0 = self._tplanes("gx", M32 * head0.K)
        db0 = self._new_dbeta(head0.K)
        th, tl = wp.wt(head4.name)
        L.conv_gemm(g4.hi, g4.lo, th, tl, head4.geom_dgrad(N), mask_hi=a0.hi, out_hi=g0.hi, out_lo=g0.lo, colsum=None if drop else db0)
        if drop:
            L.check(lib.sacb_channel_scale(L.ptr(g0.hi), L.ptr(g0.lo), L.ptr(self.dropout[0]), N, h32, w32, head0.K, st), "sacb_channel_scale")
            L.check(lib.sacb_colsum(L.ptr(g0.hi), L.ptr(g0.lo), L.ptr(db0), C.c_int64(M32), head0.K, st), "sacb_colsum")
        self._tput("gout"); self._rename("gx", "gout")
        self._wgrad(flat, wp, head0, p5, g0, grad, db0)
        gp = self.fpool.get("gp", M32 * head0.C)
        th, tl = wp.wt(head0.name)
        L.conv_gemm(g0.hi, g0.lo, th, tl, head0.geom_dgrad(N), out_f32=gp)
        self._tput("gout")
        seq = net["seq"]
        self._trunk_backward(flat, wp, seq, len(seq) - 1, None, None, gp, grad, inject={"pool30": self.inj4, "pool20": self.inj3})
        self._finalize_all(flat, wp, grad)


def make_engine(arch, N, H, W, device, train_bn=False):
    if train_bn:
        from . import engine_abn
        return engine_abn.make_engine(arch, N, H, W, device)
    return {"resnet101": ResNet101Engine, "vgg16": VGG16Engine, "fcn": FCN8sEngine}[arch](N, H, W, device)
