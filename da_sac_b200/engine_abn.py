"""Backbone schedules with TRAINING-mode batch norm -- the ABN baseline (cfg.MODEL.BASELINE = True).

The SAC path freezes every BN layer and folds it into the GEMM epilogue (engine.py).  The baseline that precedes it in the
reference's recipe (/root/reference/models/__init__.py:29 -> freeze_bn = False; train.py:113-138,281-289) trains the backbone
on the source domain with nn.SyncBatchNorm in training mode and adapts the running statistics on the target domain with
forward-only passes.  Here a conv unit becomes

    z = conv(x)              sacb_conv_gemm with a raw epilogue (same tcgen05 kernels, same weight planes)
    moments of z             sacb_bn_moments (+ one all-reduce over the ranks = SyncBatchNorm)
    mean / invstd / running  sacb_bn_train_finalize
    y = relu(bn(z) (+res))   sacb_bn_apply

and its backward   g -> sacb_bn_moments(mode 1) -> sacb_bn_bwd_finalize (d gamma, d beta) -> sacb_bn_bwd_apply (dz), after which
the filter / data gradients are the ordinary GEMMs on dz with UN-folded weights (WeightPlanes(fold_bn=False)).  Both z and y
of every unit are kept for backward (y feeds the next conv through TMA, z gives xhat).

Status: written in round 1 after the GPU budget was spent -- the kernels compile for sm_100a and the oracle for this mode is
pinned to the real reference (tests/test_abn_cpu.py), the GPU parity tests (tests/test_abn_gpu.py) have not run yet.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import engine as E
from . import lib as L

BN_MOMENTUM = 0.1          # nn.SyncBatchNorm default (the reference passes no momentum, deeplabv2.py:15,28)


class TrainBNMixin(object):
    """Buffers and per-unit building blocks of training-mode BN, shared by the three backbone schedules."""

    def _init_train_bn(self, units, n_planes):
        net, N, device = self.net, self.N, self.device
        # pre-BN conv outputs of every BN unit (kept for backward)
        self.zact = {}
        self.units = units
        kmax, ktot, pmax = 0, 0, 0
        self.bn_off = {}
        lib = L.lib()
        for s in units:
            M = N * s.hout * s.wout
            self.zact[s.name] = self._planes(M, s.K)
            self.bn_off[s.name] = ktot
            ktot += s.K
            kmax = max(kmax, s.K)
            pmax = max(pmax, int(lib.sacb_bn_moments_partial_elems(C.c_int64(M), s.K)))
        f32 = dict(device=device, dtype=torch.float32)
        f64 = dict(device=device, dtype=torch.float64)
        self.bn_mean = torch.zeros(ktot, **f32); self.bn_invstd = torch.zeros(ktot, **f32); self.bn_scale = torch.zeros(ktot, **f32)
        self.bn_partials = torch.empty(pmax, **f64)
        self.bn_sums = torch.empty(2 * kmax, **f64)          # this rank's moments
        self.bn_sums_g = torch.empty(2 * kmax, **f64)        # moments of the global batch (after the all-reduce)
        self.bn_coef = torch.empty(3 * kmax, **f32)
        # more scratch planes than the frozen-BN schedule: z of the unit in flight (no-grad passes), dz of bn3 / downsample
        self._make_pools(self.max_elems, n_planes, 2)

    # ------------------------------------------------------------------ helpers
    def _world(self):
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def _saved(self, s):
        o = self.bn_off[s.name]
        return self.bn_mean[o:o + s.K], self.bn_invstd[o:o + s.K], self.bn_scale[o:o + s.K]

    def _global_sums(self, K):
        """SyncBatchNorm: the moments of the global batch = sum of the ranks' moments (equal batch per rank, as the
        reference's loaders guarantee with drop_last); one all-reduce of 2*K doubles."""
        loc = self.bn_sums[:2 * K]
        if self._world() == 1:
            return loc
        glob = self.bn_sums_g[:2 * K]
        glob.copy_(loc)
        dist.all_reduce(glob)
        return glob

    def _bn_forward(self, flat, s, z, M, out, relu, res=None):
        lib, st = L.lib(), L.stream()
        mean, invstd, scale = self._saved(s)
        L.check(lib.sacb_bn_moments(L.ptr(z.hi), L.ptr(z.lo), None, None, None, None, 0, C.c_int64(M), s.K,
                                    L.ptr(self.bn_partials), L.ptr(self.bn_sums), st), "sacb_bn_moments")
        sums = self._global_sums(s.K)
        L.check(lib.sacb_bn_train_finalize(L.ptr(sums), C.c_double(float(M) * self._world()), L.ptr(flat.view(s.bn + ".weight")),
                                           C.c_float(E.BN_EPS), C.c_float(BN_MOMENTUM), L.ptr(flat.view(s.bn + ".running_mean")),
                                           L.ptr(flat.view(s.bn + ".running_var")), L.ptr(mean), L.ptr(invstd), L.ptr(scale),
                                           s.K, st), "sacb_bn_train_finalize")
        L.check(lib.sacb_bn_apply(L.ptr(z.hi), L.ptr(z.lo), L.ptr(mean), L.ptr(scale), L.ptr(flat.view(s.bn + ".bias")),
                                  None if res is None else L.ptr(res.hi), None if res is None else L.ptr(res.lo),
                                  1 if relu else 0, L.ptr(out.hi), L.ptr(out.lo), C.c_int64(M), s.K, st), "sacb_bn_apply")

    def _unit_train(self, flat, wp, s, xin, out, relu, keep, res=None):
        M = self.N * s.hout * s.wout
        z = self.zact[s.name] if keep else self._tplanes("z", M * s.K)
        fh, fl = wp.wf(s.name)
        L.conv_gemm(xin.hi, xin.lo, fh, fl, s.geom(self.N), out_hi=z.hi, out_lo=z.lo)        # raw conv output
        self._bn_forward(flat, s, z, M, out, relu, res)
        if not keep:
            self._tput("z")

    def _bn_backward(self, flat, grad, s, g, dz, M):
        """g: gradient at the BN output (ReLU mask already applied) -> dz (may be g itself); writes d gamma / d beta"""
        lib, st = L.lib(), L.stream()
        z = self.zact[s.name]
        mean, invstd, _ = self._saved(s)
        L.check(lib.sacb_bn_moments(L.ptr(g.hi), L.ptr(g.lo), L.ptr(z.hi), L.ptr(z.lo), L.ptr(mean), L.ptr(invstd), 1,
                                    C.c_int64(M), s.K, L.ptr(self.bn_partials), L.ptr(self.bn_sums), st), "sacb_bn_moments(bwd)")
        sums = self._global_sums(s.K)
        coef = self.bn_coef[:3 * s.K]
        L.check(lib.sacb_bn_bwd_finalize(L.ptr(self.bn_sums[:2 * s.K]), L.ptr(sums), C.c_double(float(M) * self._world()),
                                         L.ptr(flat.view(s.bn + ".weight")), L.ptr(invstd), L.ptr(grad.view(s.bn + ".weight")),
                                         L.ptr(grad.view(s.bn + ".bias")), L.ptr(coef), s.K, st), "sacb_bn_bwd_finalize")
        L.check(lib.sacb_bn_bwd_apply(L.ptr(g.hi), L.ptr(g.lo), L.ptr(z.hi), L.ptr(z.lo), L.ptr(mean), L.ptr(invstd), L.ptr(coef),
                                      L.ptr(dz.hi), L.ptr(dz.lo), C.c_int64(M), s.K, st), "sacb_bn_bwd_apply")

    def _finalize_all(self, flat, wp, grad):
        """filter gradients only: dW = sum of the split-K partials, re-laid out to OIHW (no folded-BN scale; d gamma / d beta
        were written by sacb_bn_bwd_finalize)"""
        key = (flat.buf.data_ptr(), grad.buf.data_ptr(), len(self._fin_pending), "abn")
        tab = getattr(self, "_fin_table", None)
        if tab is None or tab[0] != key:
            items, blocks = [], []
            for s, dwraw, dbeta, C_eff, RS, splits in self._fin_pending:
                # bias-only convs (VGG fc6 / fc7, the 19-class score convs): d bias = column sums of the gradient, as in the
                # frozen schedule.  BN units: d bias was zeroed by _wgrad (a bias in front of a training-mode BN has no gradient).
                bias_only = s.bn is None and s.bias and dbeta is not None
                items.append(L.FinalizeItem(L.dptr(dwraw), L.dptr(flat.view(s.name + ".weight")), None, None, None,
                                            L.dptr(dbeta) if bias_only else None,
                                            L.dptr(grad.view(s.name + ".weight")), None,
                                            L.dptr(flat.view(s.name + ".bias")) if bias_only else None,
                                            L.dptr(grad.view(s.name + ".bias")) if bias_only else None, None, s.K, C_eff, RS, splits))
                blocks.append(s.K)
            tab = self._fin_table = (key, L.item_table(items, blocks, self.device))
        items, begin, n, total = tab[1]
        L.check(L.lib().sacb_wgrad_finalize_batched(L.ptr(items), L.ptr(begin), n, total, C.c_float(E.BN_EPS), L.stream()),
                "sacb_wgrad_finalize_batched")
        self._fin_pending = []



class ResNet101TrainBNEngine(TrainBNMixin, E.ResNet101Engine):
    def __init__(self, N, H, W, device):
        E.ResNet101Engine.__init__(self, N, H, W, device)
        units = [self.net["stem"]]
        for (p, c1, c2, c3, ds) in self.net["blocks"]:
            units += [c1, c2, c3] + ([ds] if ds is not None else [])
        self._init_train_bn(units, 8)

    # ------------------------------------------------------------------ forward
    def forward(self, flat, wp, x, logits_out, keep):
        """x: fp32 NCHW [N,3,H,W]; logits_out: fp32 NCHW [N,19,h,w].  Batch statistics normalise and the running statistics
        in ``flat`` are updated (also when keep=False: the ABN target pass, train.py:281-289)."""
        net, N, lib, st = self.net, self.N, L.lib(), L.stream()
        assert not wp.fold_bn, "training-mode BN needs un-folded weight planes (WeightPlanes(fold_bn=False))"
        stem = net["stem"]
        ph, pw = net["pool_hw"]
        self.tpool_hi.reset(); self.tpool_lo.reset()
        if keep:
            self._ensure_act()
        get = (lambda tag, n: self.act[tag]) if keep else self._tplanes
        put = (lambda tag: None) if keep else self._tput
        # stem: im2col GEMM (raw) -> BN(train) -> ReLU -> max-pool
        Ms = N * stem.hout * stem.wout
        a_stem = get("stem", Ms * 64)
        zs = self.zact[stem.name] if keep else self._tplanes("z", Ms * 64)
        kp = net["stem_kp"]
        L.check(lib.sacb_stem_im2col(L.ptr(x), L.ptr(self.stem_a.hi), L.ptr(self.stem_a.lo), N, self.H, self.W,
                                     stem.hout, stem.wout, stem.R, stem.stride, stem.pad, kp, st), "sacb_stem_im2col")
        wsh, wsl = wp.stem()
        L.conv_gemm(self.stem_a.hi, self.stem_a.lo, wsh, wsl, (N, stem.hout, stem.wout, kp, stem.K, 1, 1, 1, 0),
                    out_hi=zs.hi, out_lo=zs.lo)
        self._bn_forward(flat, stem, zs, Ms, a_stem, relu=True)
        if not keep:
            self._tput("z")
        a = get("pool", N * ph * pw * 64)
        L.check(lib.sacb_maxpool_fwd(L.ptr(a_stem.hi), L.ptr(a_stem.lo), L.ptr(a.hi), L.ptr(a.lo), L.ptr(self.pool_idx),
                                     N, stem.hout, stem.wout, 64, ph, pw, st), "sacb_maxpool_fwd")
        put("stem")
        xtag = "pool"
        for (p, c1, c2, c3, ds) in net["blocks"]:
            o1 = get(c1.name, N * c1.hout * c1.wout * c1.K)
            self._unit_train(flat, wp, c1, a, o1, True, keep)
            o2 = get(c2.name, N * c2.hout * c2.wout * c2.K)
            self._unit_train(flat, wp, c2, o1, o2, True, keep)
            put(c1.name)
            if ds is not None:
                r = get(ds.name, N * ds.hout * ds.wout * ds.K)
                self._unit_train(flat, wp, ds, a, r, False, keep)
            else:
                r = a
            o3 = get(c3.name, N * c3.hout * c3.wout * c3.K)
            self._unit_train(flat, wp, c3, o2, o3, True, keep, res=r)         # relu(bn3(conv3) + residual), deeplabv2.py:94-99
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
        assert not wp.fold_bn
        self._begin_backward()
        blocks = net["blocks"]
        gout, _ = self._aspp_bwd(flat, wp, self.act[blocks[-1][3].name], dlogits, grad)
        gp = None
        for bi in range(len(blocks) - 1, -1, -1):
            (p, c1, c2, c3, ds) = blocks[bi]
            xin = self.act[blocks[bi - 1][3].name] if bi > 0 else self.act["pool"]
            o1, o2 = self.act[c1.name], self.act[c2.name]
            M = N * c3.hout * c3.wout
            # bn3: gout is also the gradient of the skip path (and of the downsample BN), so dz goes to its own planes
            gz3 = self._tplanes("gz3", M * c3.K)
            self._bn_backward(flat, grad, c3, gout, gz3, M)
            self._wgrad(flat, wp, c3, o2, gz3, grad, None)
            g2 = self._tplanes("g2", M * c2.K)
            th, tl = wp.wt(c3.name)
            L.conv_gemm(gz3.hi, gz3.lo, th, tl, c3.geom_dgrad(N), mask_hi=o2.hi, out_hi=g2.hi, out_lo=g2.lo)
            self._tput("gz3")
            # bn2 / conv2 (in place: g2 has one consumer)
            self._bn_backward(flat, grad, c2, g2, g2, M)
            self._wgrad(flat, wp, c2, o1, g2, grad, None)
            g1 = self._tplanes("g1", M * c1.K)
            th, tl = wp.wt(c2.name)
            L.conv_gemm(g2.hi, g2.lo, th, tl, c2.geom_dgrad(N), mask_hi=o1.hi, out_hi=g1.hi, out_lo=g1.lo)
            self._tput("g2")
            # bn1 / conv1
            self._bn_backward(flat, grad, c1, g1, g1, M)
            self._wgrad(flat, wp, c1, xin, g1, grad, None)
            gzd = None
            if ds is not None:
                gzd = self._tplanes("gzd", M * ds.K)
                self._bn_backward(flat, grad, ds, gout, gzd, M)
                self._wgrad(flat, wp, ds, xin, gzd, grad, None)
            Min = N * c1.hin * c1.win
            th1, tl1 = wp.wt(c1.name)
            if bi == 0:
                # block input is the max-pool output: no ReLU mask, the gradient continues through the pool
                thd, tld = wp.wt(ds.name)
                tmp = self.fpool.get("tmp", Min * c1.C)
                L.conv_gemm(gzd.hi, gzd.lo, thd, tld, ds.geom_dgrad(N), out_f32=tmp)
                gp = self.fpool.get("gpool", Min * c1.C)
                L.conv_gemm(g1.hi, g1.lo, th1, tl1, c1.geom_dgrad(N), add_f32=tmp, out_f32=gp)
                self._tput("g1"); self._tput("gout"); self._tput("gzd")
                break
            gx = self._tplanes("gx", Min * c1.C)
            if ds is None:
                L.conv_gemm(g1.hi, g1.lo, th1, tl1, c1.geom_dgrad(N), add_hi=gout.hi, add_lo=gout.lo, mask_hi=xin.hi,
                            out_hi=gx.hi, out_lo=gx.lo)
            elif c1.stride == 1:
                thd, tld = wp.wt(ds.name)
                tmp = self.fpool.get("tmp", Min * c1.C)
                L.conv_gemm(gzd.hi, gzd.lo, thd, tld, ds.geom_dgrad(N), out_f32=tmp)
                L.conv_gemm(g1.hi, g1.lo, th1, tl1, c1.geom_dgrad(N), add_f32=tmp, mask_hi=xin.hi, out_hi=gx.hi, out_lo=gx.lo)
                self.fpool.put("tmp")
            else:
                # stride-2 1x1 convs: compact data gradients on the coarse grid, scattered to the even pixels
                thd, tld = wp.wt(ds.name)
                ta = self.fpool.get("tmp", M * c1.C); tb = self.fpool.get("tmp2", M * c1.C)
                L.conv_gemm(g1.hi, g1.lo, th1, tl1, c1.geom_dgrad(N), out_f32=ta)
                L.conv_gemm(gzd.hi, gzd.lo, thd, tld, ds.geom_dgrad(N), out_f32=tb)
                L.check(lib.sacb_scatter2_mask_split(L.ptr(ta), L.ptr(tb), L.ptr(xin.hi), L.ptr(gx.hi), L.ptr(gx.lo),
                                                     N, c1.hin, c1.win, c1.C, c1.hout, c1.wout, st), "sacb_scatter2_mask_split")
                self.fpool.put("tmp"); self.fpool.put("tmp2")
            self._tput("g1"); self._tput("gout")
            if gzd is not None:
                self._tput("gzd")
            self._rename("gx", "gout")
            gout = gx
        # max-pool backward (+ ReLU mask of the stem), stem BN, stem conv
        stem = net["stem"]
        ph, pw = net["pool_hw"]
        a_stem = self.act["stem"]
        Ms = N * stem.hout * stem.wout
        gs = self._tplanes("gstem", Ms * 64)
        L.check(lib.sacb_maxpool_bwd(L.ptr(gp), L.ptr(self.pool_idx), L.ptr(a_stem.hi), L.ptr(gs.hi), L.ptr(gs.lo),
                                     N, stem.hout, stem.wout, 64, ph, pw, st), "sacb_maxpool_bwd")
        self._bn_backward(flat, grad, stem, gs, gs, Ms)
        self._first_conv_bwd(flat, wp, gs, grad, None)
        self._finalize_all(flat, wp, grad)


class _VGGTrainBN(TrainBNMixin):
    """The two VGG schedules (engine.VGG16Engine / FCN8sEngine) are chains of ``_unit`` calls forward and of
    ``_wgrad`` + data-gradient GEMM pairs backward, each gradient having exactly one consumer.  Training-mode BN therefore
    slots into those two building blocks: ``_unit`` = conv (+ bias) -> batch statistics -> normalise + ReLU, and ``_wgrad``
    first turns the gradient at the BN output into the gradient at the conv output IN PLACE, so that the parent's following
    data-gradient GEMM on the same planes already sees dz.  Everything else (pools, ASPP / FCN head wiring, dropout, score
    fusion) is the parent's verified schedule."""

    def _init_vgg_train_bn(self):
        units = [s for s in (self.net["specs"][n] for n in self.net["order"]) if s.bn is not None]
        self._init_train_bn(units, 5)
        self._flat_cur = None

    def forward(self, flat, wp, x, logits_out, keep):
        assert not wp.fold_bn, "training-mode BN needs un-folded weight planes (WeightPlanes(fold_bn=False))"
        self._flat_cur = flat
        return super().forward(flat, wp, x, logits_out, keep)

    def backward(self, flat, wp, x, dlogits, grad):
        assert not wp.fold_bn
        self._flat_cur = flat
        return super().backward(flat, wp, x, dlogits, grad)

    # ---- forward building blocks
    def _first_conv_fwd(self, flat, wp, x, out):
        lib, st, N, stem = L.lib(), L.stream(), self.N, self.net["stem"]
        kp = self.net["stem_kp"]
        L.check(lib.sacb_stem_im2col(L.ptr(x), L.ptr(self.stem_a.hi), L.ptr(self.stem_a.lo), N, self.H, self.W,
                                     stem.hout, stem.wout, stem.R, stem.stride, stem.pad, kp, st), "sacb_stem_im2col")
        sc, sh = wp.affine(stem.name)                       # fold_bn=False: scale = 1, shift = conv bias
        wsh, wsl = wp.stem()
        z = self.zact[stem.name]
        L.conv_gemm(self.stem_a.hi, self.stem_a.lo, wsh, wsl, (N, stem.hout, stem.wout, kp, stem.K, 1, 1, 1, 0),
                    scale=sc, shift=sh, out_hi=z.hi, out_lo=z.lo)
        self._bn_forward(flat, stem, z, N * stem.hout * stem.wout, out, relu=True)

    def _unit(self, wp, s, xin, out, relu, res=None):
        if s.bn is None:                                    # bias-only conv: the parent's unit (scale = 1, shift = bias)
            return super()._unit(wp, s, xin, out, relu, res)
        fh, fl = wp.wf(s.name); sc, sh = wp.affine(s.name)
        z = self.zact[s.name]
        L.conv_gemm(xin.hi, xin.lo, fh, fl, s.geom(self.N), scale=sc, shift=sh, out_hi=z.hi, out_lo=z.lo)    # z = conv + bias
        self._bn_forward(self._flat_cur, s, z, self.N * s.hout * s.wout, out, relu, res)

    # ---- backward building blocks
    def _bn_backward_unit(self, flat, grad, s, g):
        self._bn_backward(flat, grad, s, g, g, self.N * s.hout * s.wout)
        if s.bias:
            grad.view(s.name + ".bias").zero_()             # d(conv bias) = sum dz = 0 under training-mode BN

    def _wgrad(self, flat, wp, s, xin, g, grad, dbeta):
        if s.bn is not None:
            self._bn_backward_unit(flat, grad, s, g)        # g becomes dz: the caller's data-gradient GEMM reads the same planes
            dbeta = None
        super()._wgrad(flat, wp, s, xin, g, grad, dbeta)

    def _first_conv_bwd(self, flat, wp, gs, grad, dbeta):
        self._bn_backward_unit(flat, grad, self.net["stem"], gs)
        super()._first_conv_bwd(flat, wp, gs, grad, None)


class VGG16TrainBNEngine(_VGGTrainBN, E.VGG16Engine):
    def __init__(self, N, H, W, device):
        E.VGG16Engine.__init__(self, N, H, W, device)
        self._init_vgg_train_bn()


class FCN8sTrainBNEngine(_VGGTrainBN, E.FCN8sEngine):
    def __init__(self, N, H, W, device):
        E.FCN8sEngine.__init__(self, N, H, W, device)
        self._init_vgg_train_bn()


def make_engine(arch, N, H, W, device):
    return {"resnet101": ResNet101TrainBNEngine, "vgg16": VGG16TrainBNEngine, "fcn": FCN8sTrainBNEngine}[arch](N, H, W, device)
