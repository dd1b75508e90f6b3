"""``cfg.INIT_MODEL`` / ``pretrained=`` for the two VGG backbones: the reference loads a torchvision ``vgg16_bn`` snapshot into the
un-modified VGG (``vgg.load_state_dict(torch.load(pretrained))``, deeplabv2.py:248-250, fcn.py:37-39) and then re-arranges the
``features`` modules.  The B200 containers must end up with the same tensors under the reference's state_dict keys."""
import os

import pytest
import torch

torchvision = pytest.importorskip("torchvision")


@pytest.fixture(scope="module")
def snapshot(tmp_path_factory):
    torch.manual_seed(5)
    vgg = torchvision.models.vgg16_bn()
    for p in vgg.parameters():
        p.data.normal_(0, 0.05)
    for m in vgg.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    path = str(tmp_path_factory.mktemp("vgg") / "vgg16_bn.pth")
    torch.save(vgg.state_dict(), path)
    return path, vgg.state_dict()


def test_deeplab_vgg16_loads_a_torchvision_snapshot(snapshot):
    from da_sac_b200.models.deeplabv2 import DeepLabV2_VGG16
    path, sd = snapshot
    net = DeepLabV2_VGG16(19, criterion=None, pretrained=path, use_bn=True, freeze_bn=True)
    own = net.state_dict()
    # deeplabv2.py:255-260: pool4 (index 33) and pool5 (43) are dropped from ``features``; later modules move down by one
    for i in range(43):
        for suffix in ("weight", "bias", "running_mean", "running_var"):
            k = "features.%d.%s" % (i, suffix)
            if k in sd:
                j = i if i < 33 else i - 1
                assert torch.equal(own["features.%d.%s" % (j, suffix)], sd[k]), k
    assert own["features.42.weight"].shape == (1024, 512, 3, 3)           # fc6 stays a from-scratch layer


def test_fcn8s_loads_a_torchvision_snapshot(snapshot):
    from da_sac_b200.models.fcn import VGG16_FCN8s
    path, sd = snapshot
    net = VGG16_FCN8s(19, criterion=None, pretrained=path, use_bn=True, freeze_bn=True)
    own = net.state_dict()
    for k, v in sd.items():
        if not k.startswith("features."):
            continue
        i = int(k.split(".")[1])
        blk = "block1" if i < 24 else ("block2" if i < 34 else "block3")      # fcn.py:27-29
        assert torch.equal(own["%s.%d.%s" % (blk, i, k.split(".")[2])], v), k


def test_a_snapshot_of_the_wrong_architecture_is_refused(tmp_path):
    from da_sac_b200.models.fcn import VGG16_FCN8s
    path = str(tmp_path / "vgg16.pth")
    torch.save(torchvision.models.vgg16().state_dict(), path)              # no BN: keys / shapes do not match
    with pytest.raises(RuntimeError):
        VGG16_FCN8s(19, criterion=None, pretrained=path, use_bn=True, freeze_bn=True)


def test_get_model_forwards_init_model(snapshot):
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    path, sd = snapshot
    cfg = synth.ModelCfgFCN()
    cfg.INIT_MODEL = path
    net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    assert torch.equal(net.backbone.state_dict()["block2.24.weight"], sd["features.24.weight"])
    assert torch.equal(net.slow_net.state_dict()["block3.40.weight"], sd["features.40.weight"])     # the momentum copy too (models/__init__.py:38)
