"""SAC / SAC_Baseline on libsac_b200 -- the drop-in for /root/reference/models/sac.py.

Same class names, constructor, buffers (``running_conf``, ``slow_init``), ``forward`` signature and
``losses`` / ``net_outs`` contract (sac.py:315-378), so that ``train.py`` drives it unchanged:
``losses[k]`` are CUDA tensors of shape [1], ``losses["self_ce"]`` carries autograd to every student
parameter, ``y`` is mutated in place (-1 -> 255).  What differs is *how*: the teacher forward, the student
forward/backward and the whole pseudo-label tail run as a fixed schedule of sm_100a kernels; the
[BT,19,H,W] tensors of the reference (logits_up, teacher_refined, ...) are only materialised when a caller
actually reads them from ``net_outs``."""
import ctypes as C
import os

import torch
import torch.distributed as dist

from .basenet import BaseNet
from .. import engine as E
from .. import lib as L


# teacher forward + tail on a side stream (own engine), student forward on the main stream: on since round 2 (measured with the
# backward's wgrad side stream: 108.8 -> 107.5 ms per step at N=1, 110.3 -> 108.5 at N=2, profiles/r2e_*, r2g_*); SACB_TWO_STREAM=0 = off
_TWO_STREAM = os.environ.get("SACB_TWO_STREAM", "1") != "0"


class LazyOuts(dict):
    """net_outs: big diagnostic tensors are produced on first access"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._lazy = {}

    def lazy(self, key, fn):
        self._lazy[key] = fn

    def __missing__(self, key):
        if key in self._lazy:
            v = self._lazy.pop(key)()
            self[key] = v
            return v
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def keys(self):
        return list(dict.keys(self)) + list(self._lazy.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class _CELossFn(torch.autograd.Function):
    """loss_ce = CrossEntropyLoss(ignore_index=255, reduction="none")(upsample(logits), y).mean() (deeplabv2.py:217-224) through
    the fused loss kernels (``labels == NULL`` mode: plain cross-entropy against ``y``), so that the source step of the ABN
    baseline (train.py:119-138) never materialises logits_up."""

    @staticmethod
    def forward(ctx, owner, logits, y):
        ws = owner._ce_workspace(logits, y)
        L.check(L.lib().sacb_student_loss_fwd(C.byref(owner._ce_desc(logits, y, ws, 0.0, None)), L.stream()), "sacb_student_loss_fwd")
        ctx.owner, ctx.logits, ctx.y = owner, logits, y
        return ws["losses"][0:1].clone()

    @staticmethod
    def backward(ctx, g):
        dl = torch.empty_like(ctx.logits)
        ws = ctx.owner._ce_workspace(ctx.logits, ctx.y)
        L.check(L.lib().sacb_student_loss_bwd(C.byref(ctx.owner._ce_desc(ctx.logits, ctx.y, ws, 1.0, dl)), L.stream()),
                "sacb_student_loss_bwd")
        return None, dl * g, None


class SAC_Baseline(BaseNet):
    """models/sac.py:15-38: the backbone alone.  With cfg.BASELINE the backbone's BN layers train (models/__init__.py:29), which
    the B200 backbones run through engine_abn (batch statistics, running-statistics update, BN backward)."""

    def __init__(self, cfg, backbone, rank, **kwargs):
        super().__init__()
        self.backbone = backbone
        self.world_size = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = rank
        if "criterion" in kwargs:
            self.criterion = kwargs["criterion"]
        self._ce_ws = {}

    def _ce_workspace(self, logits, y):
        BT, Cn, h, w = logits.shape
        H, W = y.shape[-2:]
        key = (BT, h, w, H, W, logits.device)
        if key not in self._ce_ws:
            f32 = dict(device=logits.device, dtype=torch.float32)
            self._ce_ws[key] = dict(losses=torch.empty(2, **f32), scratch=torch.empty(2, device=logits.device, dtype=torch.float64),
                                    conf_mean=torch.zeros(H, W, **f32), running_conf=torch.zeros(Cn, **f32),
                                    grad_px=None, grad_rows=None)
        return self._ce_ws[key]

    def _ce_desc(self, logits, y, ws, grad_scale, dlogits):
        BT, Cn, h, w = logits.shape
        H, W = y.shape[-2:]
        d = L.Loss(C.sizeof(L.Loss), BT, Cn, h, w, H, W, L.ptr(logits.contiguous()), L.ptr(y), None, L.ptr(ws["conf_mean"]),
                   L.ptr(ws["running_conf"]), 3.0, L.ptr(ws["losses"]), L.ptr(ws["scratch"]), float(grad_scale), L.ptr(dlogits), None, None)
        if dlogits is not None:                         # backward: workspace of the two-stage form
            # row adjoint of the up-sampled gradient [BT, C, H, w]; the full-resolution gradient itself is never materialised
            # (csrc/sacb_tail.cu loss_grad_rows_kernel)
            if ws["grad_rows"] is None:
                ws["grad_rows"] = torch.empty(BT * Cn * H * w, device=logits.device)
                ws["grad_px"] = torch.empty(BT * Cn * H * W, device=logits.device) if W > 1216 else None    # very wide crops only
            d.grad_rows, d.grad_px = L.ptr(ws["grad_rows"]), L.ptr(ws["grad_px"])
        return d

    def forward(self, x=None, y=None, x2=None, use_teacher=False, update_teacher=False):
        bb = self.backbone
        if not hasattr(bb, "ensure_flat"):
            return bb(x, y)
        bb.ensure_flat(x.device)
        H, W = x.shape[-2:]
        from .deeplabv2 import upsample
        if y is None:                                                    # (logits, logits_up), deeplabv2.py:219-220
            with torch.no_grad():
                logits = bb.logits(x)
            return logits, upsample(logits, H, W)
        logits = bb.logits(x)
        y = y.contiguous()
        losses = {"loss_ce": _CELossFn.apply(self, logits, y)}           # deeplabv2.py:223-224
        outs = LazyOuts(logits=logits)
        outs.lazy("logits_up", lambda: upsample(logits.detach(), H, W))
        return losses, outs

    def parameter_groups(self, base_lr, wd):
        return self.backbone.parameter_groups(base_lr, wd)


class _StudentLossFn(torch.autograd.Function):
    """(loss_ce, self_ce) = fused upsample + log-softmax + NLL (deeplabv2.py:217-224, sac.py:134-149).
    Two outputs with un-materialised gradients: the backward kernel only runs for the loss that is actually
    back-propagated (self_ce on the target step, train.py:231; loss_ce on the source step, train.py:133)."""

    @staticmethod
    def forward(ctx, sac, logits, y, tail):
        desc, keep = sac._loss_desc(logits, y, tail, 0.0, None)
        L.check(L.lib().sacb_student_loss_fwd(C.byref(desc), L.stream()), "sacb_student_loss_fwd")
        ctx.sac, ctx.logits, ctx.y, ctx.tail = sac, logits, y, tail
        ctx.set_materialize_grads(False)
        # two independent tensors, not two views of one: train.py:243-245 all-reduces and divides every loss IN PLACE, which
        # autograd refuses for "the output of a function that returns multiple views"
        return keep["losses"][0:1].clone(), keep["losses"][1:2].clone()

    @staticmethod
    def backward(ctx, g_ce, g_self):
        total = None
        for g, use_labels in ((g_self, True), (g_ce, False)):
            if g is None:
                continue
            dl = torch.empty_like(ctx.logits)
            desc, keep = ctx.sac._loss_desc(ctx.logits, ctx.y, ctx.tail, 1.0, dl, use_labels=use_labels)
            L.check(L.lib().sacb_student_loss_bwd(C.byref(desc), L.stream()), "sacb_student_loss_bwd")
            dl = dl * g
            total = dl if total is None else total + dl
        return None, total, None, None


class SAC(SAC_Baseline):
    def __init__(self, cfg, backbone, slow_copy, rank, **kwargs):
        super().__init__(cfg, backbone, rank, **kwargs)
        self.cfg = cfg
        # MODEL.CONF_POOL / CONF_POOL_ON / LOSS select methods by name in the reference (sac.py:46-47,65-68,353)
        assert cfg.CONF_POOL in ("avg_pool", "minentropy_pool"), "Pooling OP _%s not found" % cfg.CONF_POOL
        assert cfg.LOSS in ("focal_ce_conf", "focal_ce"), "Pooling OP _%s not found" % cfg.LOSS
        self._pool_mode = 2 if not getattr(cfg, "CONF_POOL_ON", True) else (1 if cfg.CONF_POOL == "minentropy_pool" else 0)
        self.register_buffer("running_conf", torch.zeros(kwargs["num_classes"]))
        self.slow_net = slow_copy
        self.slow_net.eval()
        for p in self.slow_net.parameters():
            p.requires_grad = False
        self.register_buffer("slow_init", torch.Tensor([False]))
        self._engines = {}
        self._engines_teacher = {}       # only used with SACB_TWO_STREAM=1
        self._ws = {}
        self._seg = None

    # ---------------------------------------------------------------- teacher EMA (sac.py:70-102)
    def _segments(self, device):
        if self._seg is None or self._seg[0].device != device:
            f = self.backbone._flat
            r = torch.tensor(f.ranges(), dtype=torch.int64, device=device)
            self._seg = (r, len(f.entries), torch.zeros(len(f.entries), device=device), torch.zeros(1, device=device))
        return self._seg

    @torch.no_grad()
    def _momentum_update(self, update=False):
        dev = self.running_conf.device
        fs, ft = self.backbone.ensure_flat(dev), self.slow_net.ensure_flat(dev)
        # the reference reads slow_init on the host every call (sac.py:75); while a CUDA graph is being captured the
        # teacher is known to be initialised (capture only happens after eager warm-up steps)
        initialised = True if torch.cuda.is_current_stream_capturing() else bool(self.slow_init[0])
        if not initialised:
            self.running_conf.fill_(self.cfg.THRESHOLD_BETA)
            self.slow_init[0] = True
            ft.buf.copy_(fs.buf)                        # slow_net.load_state_dict(backbone.state_dict())
            self.slow_net.mark_dirty()
            return torch.zeros(1, device=dev)
        ranges, nseg, seg_sq, out = self._segments(dev)
        L.check(L.lib().sacb_ema_norm(L.ptr(ft.buf), L.ptr(fs.buf), L.ptr(ranges), nseg, C.c_float(self.cfg.NET_MOMENTUM),
                                      1 if update else 0, L.ptr(seg_sq), L.ptr(out), L.stream()), "sacb_ema_norm")
        if update:
            self.slow_net.mark_dirty()
        return out.clone()

    # ---------------------------------------------------------------- tail
    def _workspace(self, BT, T, H, W, dev):
        key = (BT, T, H, W)
        if key not in self._ws:
            lib = L.lib()
            Cn = E.NUM_CLASSES
            f32 = dict(device=dev, dtype=torch.float32)
            self._ws[key] = dict(
                probs=torch.empty(lib.sacb_tail_probs_elems(BT, Cn, H, W), **f32),
                pooled=torch.empty(lib.sacb_tail_pooled_elems(BT // T, Cn, H, W), **f32),
                part_sums=torch.empty(lib.sacb_tail_part_sums_elems(BT, Cn, H, W), **f32),
                peaks=torch.empty(BT * Cn, **f32), thresholds=torch.empty(BT, Cn, **f32),
                conf=torch.empty(BT, 1, H, W, **f32), idx=torch.empty(BT, 1, H, W, device=dev, dtype=torch.uint8),
                labels=torch.empty(BT, H, W, device=dev, dtype=torch.uint8), conf_mean=torch.empty(H, W, **f32),
                losses=torch.empty(2, **f32), scratch=torch.empty(2, device=dev, dtype=torch.float64))
        return self._ws[key]

    def _tail(self, teacher_logits, y_raw, affine, affine_inv, T, refined=None, refine_only=False):
        BT, Cn, h, w = teacher_logits.shape
        H, W = y_raw.shape[-2:]
        # fractional group (sac.py:243-245, _gather :198-216): this rank holds T0 < T views of ONE group; the other
        # views live on the neighbouring ranks (train.py:185-209).  The kernels then run with T = T0 and the
        # reference-frame partial sums are exchanged between the ranks sharing the group.
        T0 = min(T, BT)
        ws = self._workspace(BT, T0, H, W, teacher_logits.device)
        cfg = self.cfg
        A, Ai = affine.contiguous().float(), affine_inv.contiguous().float()

        def desc(phase):
            return L.Tail(C.sizeof(L.Tail), BT, T0, Cn, h, w, H, W, L.ptr(teacher_logits), L.ptr(y_raw),
                          L.ptr(A), L.ptr(Ai), L.ptr(self.running_conf),
                          1 if self.training else 0, 1 if cfg.CONF_DISCOUNT else 0,
                          cfg.THRESHOLD_BETA, cfg.STAT_MOMENTUM, cfg.RUN_CONF_UPPER, cfg.RUN_CONF_LOWER,
                          L.ptr(ws["probs"]), L.ptr(ws["pooled"]), L.ptr(ws["part_sums"]), L.ptr(ws["peaks"]),
                          L.ptr(ws["conf"]), L.ptr(ws["idx"]), L.ptr(ws["labels"]), L.ptr(ws["conf_mean"]),
                          L.ptr(ws["thresholds"]), L.ptr(refined), phase, self._pool_mode)
        if refine_only:
            # lazy net_outs["teacher_refined"]: one more warp of the pooled probabilities this forward left in the workspace;
            # no running_conf update, no label rewrite, no exchange between ranks
            L.check(L.lib().sacb_teacher_tail(C.byref(desc(3)), L.stream()), "sacb_teacher_tail(refined)")
            return ws
        if T0 == T or self._pool_mode == 2:
            L.check(L.lib().sacb_teacher_tail(C.byref(desc(0)), L.stream()), "sacb_teacher_tail")
            return ws
        assert self._pool_mode == 0, "fractional view-groups are implemented for CONF_POOL=avg_pool (the reference's _gather lives there)"
        d1 = desc(1)
        L.check(L.lib().sacb_teacher_tail(C.byref(d1), L.stream()), "sacb_teacher_tail(partial sums)")
        self._exchange_partial_sums(ws["pooled"], BT, T)
        d2 = desc(2)
        L.check(L.lib().sacb_teacher_tail(C.byref(d2), L.stream()), "sacb_teacher_tail(labels)")
        return ws

    def _exchange_partial_sums(self, pooled, B, T):
        """Sum the reference-frame partial sums over the ranks that share this rank's view-group.  The reference
        all-gathers every rank's [B,19,H,W] probabilities and concatenates ``stride`` of them (sac.py:204-214); only
        their sum over views is ever used (sac.py:252-253), so one sum all-reduce inside the sub-group is enough."""
        from ..trainer import fractional_subgroup
        first, stride = fractional_subgroup(self.rank, B, T)
        assert dist.is_initialized() and dist.get_world_size() >= first + stride, \
            "fractional view-groups (local batch %d < GROUP_SIZE %d) need torch.distributed with >= %d ranks" % (B, T, first + stride)
        key = (dist.get_world_size(), stride)
        if getattr(self, "_subgroups", None) is None or self._subgroups[0] != key:
            # new_group is collective: every rank creates every sub-group, in the same order
            groups = {}
            for f in range(0, dist.get_world_size() - stride + 1, stride):
                groups[f] = dist.new_group(list(range(f, f + stride)))
            self._subgroups = (key, groups)
        dist.all_reduce(pooled, group=self._subgroups[1][first])

    def _loss_desc(self, logits, y, tail, grad_scale, dlogits, use_labels=True):
        BT, Cn, h, w = logits.shape
        H, W = y.shape[-2:]
        keep = tail
        conf_mean = tail["conf_mean"]
        if self.cfg.LOSS == "focal_ce":                 # sac.py:119-132: no confidence weighting
            if "ones" not in tail:
                tail["ones"] = torch.ones_like(conf_mean)
            conf_mean = tail["ones"]
        d = L.Loss(C.sizeof(L.Loss), BT, Cn, h, w, H, W, L.ptr(logits.contiguous()), L.ptr(y), L.ptr(tail["labels"]) if use_labels else None,
                   L.ptr(conf_mean), L.ptr(self.running_conf), float(self.cfg.FOCAL_P), L.ptr(tail["losses"]),
                   L.ptr(tail["scratch"]), float(grad_scale), L.ptr(dlogits), None, None)
        if dlogits is not None:                         # backward: workspace of the two-stage form
            if "grad_rows" not in tail:
                tail["grad_rows"] = torch.empty(BT * Cn * H * w, device=logits.device)
                tail["grad_px"] = torch.empty(BT * Cn * H * W, device=logits.device) if W > 1216 else None  # very wide crops only
            d.grad_rows, d.grad_px = L.ptr(tail["grad_rows"]), L.ptr(tail["grad_px"])
        return d, keep

    # ---------------------------------------------------------------- forward (sac.py:315-378)
    def forward(self, x, y=None, x2=None, affine=None, affine_inv=None,
                use_teacher=False, update_teacher=False, reset_teacher=False, T=None, teacher=False):
        dev = x.device
        self.backbone.ensure_flat(dev); self.slow_net.ensure_flat(dev)
        if y is None:                                                    # inference-only mode (sac.py:324-329)
            net = self.slow_net if teacher else self.backbone
            with torch.no_grad():
                logits = net.logits(x, self._engines, refresh=True)
            from .deeplabv2 import upsample
            return logits, upsample(logits, *x.shape[-2:])
        if reset_teacher:
            self.slow_init[0] = False
        # the kernels read raw device memory: labels as int64 (what the loaders produce and nn.CrossEntropyLoss demands of the
        # reference), images / affine matrices as fp32 (converted below where needed)
        if y.dtype != torch.int64:
            raise TypeError("SAC.forward: y must be int64 (got %s)" % y.dtype)
        if use_teacher and (x2 is None or affine is None or affine_inv is None or T is None):
            raise ValueError("SAC.forward(use_teacher=True) needs x2, affine, affine_inv and T")
        y_raw = y.clone()                                                # keeps the -1 padding marker for the kernels
        y.masked_fill_(y == -1, 255)                                     # in-place on the caller's tensor (sac.py:337-338)
        losses = {}
        H, W = x.shape[-2:]
        if update_teacher:
            losses["teacher_diff"] = self._momentum_update(True)         # sac.py:342-344
        tail = None
        # SACB_TWO_STREAM (default on): the teacher forward + tail and the
        # student forward are independent until the loss, so they are issued on two streams.  Every GEMM is a persistent
        # kernel that owns all SMs, but its last partial wave (5.36 waves on the 256-channel layers) and the launch gaps leave
        # SMs idle that the other stream's next kernel can take.  The teacher then needs its own engine (im2col matrix, ASPP
        # scratch and plane pools are per engine).
        two_stream = use_teacher and _TWO_STREAM and not L.SERIALIZE and L.on_device(x)
        side = None
        if use_teacher:
            self.slow_net.eval()
            if two_stream:
                side = self._side_stream = getattr(self, "_side_stream", None) or torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())           # EMA / inputs are ordered before the fork
                with torch.cuda.stream(side), torch.no_grad():
                    t_logits = self.slow_net.logits(x2, self._engines_teacher, refresh=False)
                    tail = self._tail(t_logits, y_raw, affine, affine_inv, T)
            else:
                with torch.no_grad():                                    # sac.py:348-350
                    t_logits = self.slow_net.logits(x2, self._engines, refresh=False)
                tail = self._tail(t_logits, y_raw, affine, affine_inv, T)   # sac.py:353-357
        s_logits = self.backbone.logits(x, self._engines, refresh=True)  # sac.py:340
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)               # join: the loss reads the pseudo labels
        if tail is None:
            BT = x.shape[0]
            tail = self._workspace(BT, 1, H, W, dev)
            tail["labels"].fill_(255); tail["conf_mean"].zero_()
        loss_ce, self_ce = _StudentLossFn.apply(self, s_logits, y_raw, tail)
        losses["loss_ce"] = loss_ce
        outs = LazyOuts(logits=s_logits)
        from .deeplabv2 import upsample
        outs.lazy("logits_up", lambda: upsample(s_logits.detach(), H, W))
        if use_teacher:
            losses["self_ce"] = self_ce                                  # sac.py:360-361
            # the tail's buffers are workspace that the next forward overwrites; what net_outs hands out are copies, made when
            # (and if) a caller reads them -- the reference returns fresh tensors (sac.py:362-371)
            outs.lazy("teacher_conf", lambda: tail["conf"].clone())
            outs["running_conf"] = self.running_conf
            outs.lazy("teacher_labels", lambda: tail["labels"].long())
            outs.lazy("teacher_init", lambda: upsample(t_logits, H, W))

            def _refined():
                r = torch.empty(x.shape[0], E.NUM_CLASSES, H, W, device=dev)
                self._tail(t_logits, y_raw, affine, affine_inv, T, refined=r, refine_only=True)
                return r
            outs.lazy("teacher_refined", _refined)
            if self._pool_mode != 2:
                # diagnostics of _refine (sac.py:292-296; read by base_trainer._visualise :171-173 only): the masked teacher
                # probabilities and the clean frames warped to the reference frame.  Visualisation path, not the hot path: plain
                # torch ops on the up-sampled teacher logits, produced on first access.
                def _aligned(t):
                    grid = torch.nn.functional.affine_grid(affine.float(), size=t.size(), align_corners=False)
                    return torch.nn.functional.grid_sample(t, grid, align_corners=False)

                def _teacher_aligned():
                    p = torch.softmax(upsample(t_logits, H, W), 1)
                    p *= 1 - (y_raw == -1)[:, None].type_as(p)
                    return _aligned(p)
                outs.lazy("teacher_aligned", _teacher_aligned)
                outs.lazy("frames_aligned", lambda: _aligned(x2.float()))
            losses["teacher_diff"] = self._momentum_update(False)        # sac.py:374
        return losses, outs

    def parameter_groups(self, base_lr, wd):
        return self.backbone.parameter_groups(base_lr, wd)
