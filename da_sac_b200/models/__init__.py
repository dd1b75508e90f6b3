"""Model registry of the B200 drop-in: ``from models import get_model`` keeps working after
``sys.modules["models"] = da_sac_b200.models``.

Contract taken from /root/reference/models/__init__.py:14-41 and train.py:88-89: ``get_model(cfg.MODEL, rank,
num_classes=19, criterion=...)`` builds the backbone named by ``cfg.ARCH`` -- twice when the consistency loss is trained,
the second copy being the momentum (teacher) network -- and wraps it in ``SAC``; with ``cfg.BASELINE`` it returns
``SAC_Baseline`` around a single backbone whose BN layers train (the ABN baseline).  ``cfg.INIT_MODEL`` is forwarded as
``pretrained`` when the file exists."""
import os

from .deeplabv2 import DeepLabV2_ResNet101, DeepLabV2_VGG16
from .fcn import VGG16_FCN8s
from .sac import SAC, SAC_Baseline

# ARCH key -> (backbone class, fixed constructor arguments)
BACKBONES = {
    "deeplabv2_resnet101": (DeepLabV2_ResNet101, {}),
    "deeplabv2_vgg16_bn": (DeepLabV2_VGG16, {"use_bn": True}),
    "fcn_vgg16_bn": (VGG16_FCN8s, {"use_bn": True}),
}


def _build_backbone(arch, args, kwargs):
    cls, fixed = BACKBONES[arch]
    return cls(*args, **dict(kwargs, **fixed))


def get_model(cfg, rank, *args, **kwargs):
    arch = cfg.ARCH.lower()
    if arch not in BACKBONES:
        raise NotImplementedError("libsac_b200 has no backbone '%s'; available: %s" % (arch, sorted(BACKBONES)))
    opts = dict(kwargs)
    snapshot = cfg.INIT_MODEL
    if snapshot and os.path.isfile(snapshot):
        opts["pretrained"] = snapshot
    # the SAC stage trains with frozen BN statistics; only the ABN baseline lets them move (reference: models/__init__.py:29)
    opts["freeze_bn"] = not cfg.BASELINE
    student = _build_backbone(arch, args, opts)
    if cfg.BASELINE:
        return SAC_Baseline(cfg, student, rank, **opts)
    teacher = _build_backbone(arch, args, opts)
    return SAC(cfg, student, teacher, rank, **opts)
