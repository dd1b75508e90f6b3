"""DeepLabV2_ResNet101 on libsac_b200.

Same constructor arguments, ``forward(im, y=None)`` contract, ``state_dict`` key layout
(``model.conv1``, ``model.layer3.5.bn2`` ... ``model.layer5.conv2d_list.3``) and optimiser
groups as /root/reference/models/deeplabv2.py:173-227, but the nn.Conv2d / nn.BatchNorm2d
modules below are parameter containers only: compute runs in ``engine.ResNet101Engine``
(tcgen05 implicit-GEMM kernels), wired into autograd by ``_BackboneFn`` so that DDP gradient
hooks and ``torch.optim`` keep working on the real nn.Parameters."""
import torch
import torch.nn as nn

from .basenet import BaseNet
from .. import engine as E
from .. import lib as L


def _bn(c):
    return nn.BatchNorm2d(c, eps=E.BN_EPS)


class _Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride, dilation, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = _bn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _bn(planes * 4)
        self.downsample = downsample


class _Classifier(nn.Module):
    def __init__(self, fan_in, dilations, num_classes):
        super().__init__()
        self.conv2d_list = nn.ModuleList(
            [nn.Conv2d(fan_in, num_classes, 3, stride=1, padding=d, dilation=d, bias=True) for d in dilations])
        for m in self.conv2d_list:
            m.weight.data.normal_(0, 0.01)


class _ResNetParams(nn.Module):
    def __init__(self, layers, num_classes):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = _bn(64)
        inplanes = 64
        for li, (planes, n, stride, dil) in enumerate(((64, layers[0], 1, 1), (128, layers[1], 2, 1),
                                                       (256, layers[2], 1, 2), (512, layers[3], 1, 4)), 1):
            blocks = []
            for b in range(n):
                ds = None
                if b == 0:
                    ds = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False), _bn(planes * 4))
                blocks.append(_Bottleneck(inplanes, planes, stride if b == 0 else 1, dil, ds))
                inplanes = planes * 4
            setattr(self, "layer%d" % li, nn.Sequential(*blocks))
        self.layer5 = _Classifier(2048, (6, 12, 18, 24), num_classes)
        for m in self.modules():                      # stock init of the reference (deeplabv2.py:135-141)
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, 0.01)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1); m.bias.data.zero_()


class _BackboneFn(torch.autograd.Function):
    """logits = ResNet101-DeepLabv2(x); all parameter gradients come from engine.backward"""

    @staticmethod
    def forward(ctx, net, eng, x, *params):
        logits = torch.empty(x.shape[0], E.NUM_CLASSES, *eng.net["out_hw"], device=x.device)
        net._pre_forward(eng, x, True)
        eng.forward(net._flat, net._planes(True), x, logits, keep=True)
        ctx.net, ctx.eng, ctx.x = net, eng, x
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        net, eng = ctx.net, ctx.eng
        net._select_grad_buffer()
        L.set_phase("bwd")                       # precision policy: the gradient GEMMs may run single-pass bf16 (lib.PRECISION)
        try:
            eng.backward(net._flat, net._planes(True), ctx.x, dlogits.contiguous(), net._grad)
        finally:
            L.set_phase("fwd")
        return (None, None, None) + tuple(net._grad.view(k) for k in net._param_keys)


class _EngineBackbone(BaseNet):
    """Shared machinery of the B200 backbones: flat fp32 parameter storage, bf16 weight planes, per-shape engines."""
    ARCH = None

    def _init_engine_state(self, freeze_bn=True):
        # freeze_bn=False is the ABN baseline (models/__init__.py:29): BN layers normalise with batch statistics and update
        # their running statistics while the module is in train() mode (engine_abn.py); eval() uses the frozen-BN engine
        self._train_bn = not freeze_bn
        self._flat = None
        self._grad = None
        self._wp = None
        self._wp_version = -1
        self._version = 0            # bumped whenever parameter values change
        self._engines = {}

    def lr_mult(self):
        return 1., 10.

    def lr_mult_bias(self):
        return 2., 20.

    # ---------------------------------------------------------------- flat parameter storage
    def _named_tensors(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        return sd

    def ensure_flat(self, device):
        """(Re)home every parameter / BN statistic in one contiguous fp32 buffer (needed by the multi-tensor
        EMA / SGD kernels and by the weight-plane preparation). Parameter objects keep their identity."""
        tensors = self._named_tensors()
        if self._flat is not None and self._flat.buf.device == device:
            ok = all(tensors[k].data_ptr() == self._flat.view(k).data_ptr() for k, _, _ in self._flat.entries)
            if ok:
                return self._flat
        net = E.build_net(self.ARCH, 64, 64)
        flat = E.FlatParams(net, device)
        for k, _, _ in flat.entries:
            v = flat.view(k)
            v.copy_(tensors[k].detach())
            tensors[k].data = v
        self._flat = flat
        self._grad = E.FlatParams(net, device)
        self._param_keys = [k for k, _, is_p in flat.entries if is_p]
        self._params = [tensors[k] for k in self._param_keys]
        self._wp = None
        self._version += 1
        return flat

    def mark_dirty(self):
        self._version += 1

    def _select_grad_buffer(self):
        """The engine WRITES a whole backward pass into the flat buffer ``_grad`` and autograd receives views of it, so
        after the first backward ``p.grad`` aliases that buffer.  A second backward without ``zero_grad`` in between
        (train.py:128-138 then :231-232: the source pass and the target pass accumulate into one optimiser step) must
        therefore not overwrite it: switch to the alternate flat buffer, autograd then adds the two."""
        g = self._grad
        lo, hi = g.buf.data_ptr(), g.buf.data_ptr() + g.buf.numel() * 4
        held = getattr(self, "_grad_held", False) or \
            any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self._params)
        if held:
            if getattr(self, "_grad_alt", None) is None:
                self._grad_alt = E.FlatParams(E.build_net(self.ARCH, 64, 64), g.buf.device)
            self._grad, self._grad_alt = self._grad_alt, self._grad
            self._grad_held = False

    def hold_grad(self):
        """Keep the flat gradient buffer of the backward pass that just ran (the next backward goes to the alternate
        buffer) and return it: the joint source+target step adds the two with ONE flat kernel instead of letting
        autograd accumulate 320 tensors one by one (trainer.JointStepper)."""
        self._grad_held = True
        return self._grad

    def _pre_forward(self, eng, x, with_grad):
        """hook for per-forward engine state (FCN dropout masks)"""
        return None

    def _bn_training(self):
        return self._train_bn and self.training

    def _planes(self, with_dgrad):
        fold = not self._bn_training()
        if self._wp is None or self._wp.with_dgrad != with_dgrad or self._wp.fold_bn != fold:
            self._wp = E.WeightPlanes(E.build_net(self.ARCH, 64, 64), self._flat.buf.device, with_dgrad, fold_bn=fold)
            self._wp_version = -1
        if self._wp_version != self._version:
            self._wp.prepare(self._flat)
            self._wp_version = self._version
        return self._wp

    def engine(self, N, H, W, shared=None):
        train_bn = self._bn_training()
        key = (self.ARCH, N, H, W) + (("train_bn",) if train_bn else ())
        cache = self._engines if shared is None else shared
        if key not in cache:
            cache[key] = E.make_engine(self.ARCH, N, H, W, self._flat.buf.device, train_bn=train_bn)
        return cache[key]

    def _after_forward(self):
        """training-mode BN also counts its batches (nn.BatchNorm2d.num_batches_tracked; unused by momentum=0.1 BN, kept for
        checkpoint parity)"""
        if self._bn_training():
            nbt = [m.num_batches_tracked for m in self.modules() if isinstance(m, nn.BatchNorm2d)]
            torch._foreach_add_(nbt, 1)

    # ---------------------------------------------------------------- forward
    def logits(self, im, engines=None, refresh=True):
        """student (autograd) or teacher / inference (no-grad) forward. ``refresh`` re-derives the bf16 weight
        planes and folded BN affine from the fp32 parameters (needed whenever an optimiser or EMA changed them)."""
        self.ensure_flat(im.device)
        im = im.contiguous().float()
        eng = self.engine(im.shape[0], im.shape[2], im.shape[3], engines)
        trainable = any(p.requires_grad for p in self._params)
        if refresh:
            self.mark_dirty()
        if trainable and torch.is_grad_enabled():
            out = _BackboneFn.apply(self, eng, im, *self._params)
            self._after_forward()
            return out
        out = torch.empty(im.shape[0], E.NUM_CLASSES, *eng.net["out_hw"], device=im.device)
        self._pre_forward(eng, im, False)
        eng.forward(self._flat, self._planes(trainable), im, out, keep=False)
        self._after_forward()
        return out

    def forward(self, im, y=None):
        """(logits, logits_up) for inference, (losses, outs) with a label map (deeplabv2.py:213-227)."""
        logits = self.logits(im)
        H, W = im.shape[-2:]
        up = upsample(logits.detach() if not logits.requires_grad else logits, H, W)
        if y is None:
            return logits, up
        ce = self.criterion(up, y)
        return {"loss_ce": ce.mean().view(1)}, {"logits_up": up, "logits": logits}


class DeepLabV2_ResNet101(_EngineBackbone):
    """/root/reference/models/deeplabv2.py:173-227"""
    ARCH = "resnet101"

    def __init__(self, num_classes=20, criterion=nn.CrossEntropyLoss(ignore_index=255, reduction="none"),
                 pretrained=None, freeze_bn=False):
        super().__init__()
        assert num_classes == E.NUM_CLASSES, "libsac_b200 kernels are built for 19 classes"
        self.model = _ResNetParams([3, 4, 23, 3], num_classes)
        if pretrained is not None:
            self.model.load_state_dict(torch.load(pretrained), strict=False)
        if freeze_bn:
            self._freeze_bn(self)
        self._from_scratch(self.model.layer5)
        self.criterion = criterion
        self._init_engine_state(freeze_bn)


class DeepLabV2_VGG16(_EngineBackbone):
    """/root/reference/models/deeplabv2.py:229-312 (use_bn=True): torchvision vgg16_bn features with conv5 dilated,
    pool4/pool5 removed, fc6/fc7 as dilated 3x3 convs, ASPP on 1024 channels. Module indices inside ``features`` match
    the reference so that checkpoints (``features.N.*``, ``classifier.conv2d_list.*``) load unchanged."""
    ARCH = "vgg16"

    def __init__(self, num_classes, criterion=None, pretrained=None, use_bn=False, freeze_bn=False):
        super().__init__()
        assert use_bn, "libsac_b200 implements the BN variant used by the reference configs (deeplabv2_vgg16_bn)"
        assert num_classes == E.NUM_CLASSES, "libsac_b200 kernels are built for 19 classes"
        self.criterion = criterion
        layers = []
        pool_after = set(E.VGG16_POOL_AFTER)
        for (idx, cin, cout, dil) in E.VGG16_CONVS:
            assert len(layers) == idx
            layers += [nn.Conv2d(cin, cout, 3, padding=dil, dilation=dil), _bn(cout), nn.ReLU(inplace=True)]
            if idx in pool_after:
                layers.append(nn.MaxPool2d(2, 2))
        fc6 = nn.Conv2d(512, 1024, 3, padding=4, dilation=4)
        fc7 = nn.Conv2d(1024, 1024, 3, padding=4, dilation=4)
        layers += [fc6, nn.ReLU(inplace=True), fc7, nn.ReLU(inplace=True)]
        self.features = nn.Sequential(*layers)
        if pretrained is not None:
            # deeplabv2.py:248-250 loads a torchvision vgg16_bn snapshot into the un-modified VGG and then drops pool4 (index 33)
            # and pool5 (43) from ``features``, so the conv5 modules 34 / 37 / 40 (+ their BN) move down by one
            print("VGG16: Loading snapshot: ", pretrained)
            load_torchvision_vgg16_bn(torch.load(pretrained), {"features": self.features}, lambda i: ("features", i if i < 33 else i - 1))
        self.classifier = _Classifier(1024, (6, 12, 18, 24), num_classes)
        if freeze_bn:
            self._freeze_bn(self)
        self._from_scratch(self.classifier)
        self._from_scratch(fc6)
        self._from_scratch(fc7)
        self._init_engine_state(freeze_bn)


def load_torchvision_vgg16_bn(sd, containers, where):
    """Copy the ``features.N.*`` tensors of a torchvision ``vgg16_bn`` state dict (what ``vgg.load_state_dict(torch.load(
    pretrained))`` consumes in deeplabv2.py:249 / fcn.py:39) into the parameter containers of a B200 backbone.
    ``where(N) -> (container name, index inside it)``.  Like the reference's strict load, every ``features`` tensor must find
    its place and every conv / BN of the trunk must be covered; the ``classifier.*`` (fc) entries of the snapshot are required to
    exist as torchvision writes them, and are dropped exactly as the reference drops ``vgg.classifier``."""
    want = {}
    for key, v in sd.items():
        parts = key.split(".")
        if parts[0] == "classifier":
            continue
        if parts[0] != "features" or len(parts) != 3:
            raise RuntimeError("unexpected key in vgg16_bn snapshot: %s" % key)
        name, idx = where(int(parts[1]))
        want["%s.%d.%s" % (name, idx, parts[2])] = v
    have = {}
    for name, seq in containers.items():
        for k, t in seq.state_dict().items():
            have["%s.%s" % (name, k)] = t
    trunk = {k for k in have if int(k.split(".")[1]) <= 41 or k.split(".")[0] != "features"}    # fc6 / fc7 (42, 44) are new layers
    missing = sorted(k for k in trunk if k not in want)
    extra = sorted(k for k in want if k not in have)
    if missing or extra:
        raise RuntimeError("vgg16_bn snapshot does not match: missing %s, unexpected %s" % (missing[:4], extra[:4]))
    with torch.no_grad():
        for k in trunk:
            if have[k].shape != want[k].shape:
                raise RuntimeError("size mismatch for %s: %s vs %s" % (k, tuple(want[k].shape), tuple(have[k].shape)))
            have[k].copy_(want[k])


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        B, Cc, h, w = x.shape
        out = torch.empty(B, Cc, H, W, device=x.device)
        L.check(L.lib().sacb_upsample(L.ptr(x.contiguous()), L.ptr(out), B, Cc, h, w, H, W, L.stream()), "sacb_upsample")
        return out

    @staticmethod
    def backward(ctx, g):
        raise L.SacbError("gradient through the materialised logits_up is not on the B200 hot path; "
                          "use SAC.forward's fused student loss (losses['self_ce'])")


def upsample(x, H, W):
    """F.interpolate(x, (H, W), mode='bilinear', align_corners=True) (deeplabv2.py:217)"""
    return _UpsampleFn.apply(x, H, W)
