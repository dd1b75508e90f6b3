"""Model registry -- drop-in for /root/reference/models/__init__.py:14-41.

``get_model(cfg.MODEL, rank, num_classes=19, criterion=...)`` returns ``SAC`` (or ``SAC_Baseline`` when
``cfg.BASELINE``) wrapping a B200-native backbone and its momentum copy."""
import os

from functools import partial

from .deeplabv2 import DeepLabV2_ResNet101, DeepLabV2_VGG16
from .fcn import VGG16_FCN8s
from .sac import SAC, SAC_Baseline


def get_model(cfg, rank, *args, **kwargs):
    models = {"deeplabv2_resnet101": DeepLabV2_ResNet101, "deeplabv2_vgg16_bn": partial(DeepLabV2_VGG16, use_bn=True),
              "fcn_vgg16_bn": partial(VGG16_FCN8s, use_bn=True)}
    arch = cfg.ARCH.lower()
    if arch not in models:
        raise NotImplementedError("libsac_b200: backbone '%s' is not built yet (SURVEY.md section 8(f)); available: %s"
                                  % (arch, sorted(models)))
    if len(cfg.INIT_MODEL) > 0 and os.path.isfile(cfg.INIT_MODEL):
        kwargs["pretrained"] = cfg.INIT_MODEL
    kwargs["freeze_bn"] = not cfg.BASELINE
    backbone = models[arch](*args, **kwargs)
    if cfg.BASELINE:
        return SAC_Baseline(cfg, backbone, rank, **kwargs)
    slow_copy = models[arch](*args, **kwargs)
    return SAC(cfg, backbone, slow_copy, rank, **kwargs)
