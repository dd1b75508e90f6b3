"""VGG16_FCN8s on libsac_b200 -- drop-in for /root/reference/models/fcn.py:10-149 (use_bn=True variant).

Same constructor, ``forward(x, y=None)`` contract, optimiser groups and ``state_dict`` keys (``block1.N``, ``block2.N``,
``block3.N`` with torchvision's indices, ``vgg_head.{0,1,4,5,8}``, ``score_pool4``, ``score_pool3``); the modules are
parameter containers, compute runs in ``engine.FCN8sEngine``."""
from collections import OrderedDict

import torch
import torch.nn as nn

from .deeplabv2 import _EngineBackbone, _bn, load_torchvision_vgg16_bn
from .. import engine as E


class VGG16_FCN8s(_EngineBackbone):
    ARCH = "fcn"

    def __init__(self, num_classes, criterion=None, pretrained=None, use_bn=False, freeze_bn=False, drop_rate=0.1):
        super().__init__()
        assert use_bn, "libsac_b200 implements the BN variant used by the reference configs (fcn_vgg16_bn)"
        assert num_classes == E.NUM_CLASSES, "libsac_b200 kernels are built for 19 classes"
        self.criterion = criterion
        self.drop_rate = drop_rate
        for blk, convs, pools in E.FCN_BLOCKS:
            mods = OrderedDict()
            for (idx, cin, cout) in convs:
                mods[str(idx)] = nn.Conv2d(cin, cout, 3, padding=1)
                mods[str(idx + 1)] = _bn(cout)
                mods[str(idx + 2)] = nn.ReLU(inplace=True)
                if idx in pools:
                    mods[str(idx + 3)] = nn.MaxPool2d(2, 2)
            setattr(self, blk, nn.Sequential(mods))
        if pretrained is not None:
            # fcn.py:27-39: block1 / block2 / block3 are slices [:24] / [24:34] / [34:] of vgg.features and keep torchvision's
            # module indices, so ``features.N`` lands in the block that holds index N
            print("VGG16-FCN8s: Loading snapshot: ", pretrained)
            load_torchvision_vgg16_bn(torch.load(pretrained), {b: getattr(self, b) for b in ("block1", "block2", "block3")},
                                      lambda i: ("block1" if i < 24 else ("block2" if i < 34 else "block3"), i))
        else:
            print("VGG16-FCN8s: Initialising from scratch")
        self.vgg_head = nn.Sequential(
            nn.Conv2d(512, 4096, 7, padding=3), _bn(4096), nn.ReLU(inplace=True), nn.Dropout2d(p=drop_rate),
            nn.Conv2d(4096, 4096, 1), _bn(4096), nn.ReLU(inplace=True), nn.Dropout2d(p=drop_rate),
            nn.Conv2d(4096, num_classes, 1))
        if freeze_bn:
            self._freeze_bn(self)
        self._from_scratch(self.vgg_head)
        self.score_pool4 = nn.Conv2d(512, num_classes, 1)
        self.score_pool4.weight.data.normal_(0, 0.01)
        self._from_scratch(self.score_pool4)
        self.score_pool3 = nn.Conv2d(256, num_classes, 1)
        self.score_pool3.weight.data.normal_(0, 0.01)
        self._from_scratch(self.score_pool3)
        self._init_engine_state(freeze_bn)

    def _pre_forward(self, eng, x, with_grad):
        """Dropout2d(p) of the head (fcn.py:52,56): active in train mode only; one Bernoulli draw per (sample, channel)"""
        if with_grad and self.training and self.drop_rate > 0:
            keep = 1.0 - self.drop_rate
            eng.dropout = tuple((torch.rand(x.shape[0], 4096, device=x.device) < keep).float() / keep for _ in range(2))
        else:
            eng.dropout = None
