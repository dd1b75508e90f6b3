"""Golden vectors for the ABN baseline mode (cfg.MODEL.BASELINE = True), produced by the REAL reference on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_abn.py

Writes tests/golden/abn_resnet101_tiny.npz.  Protocol = one iteration of train_epoch in BASELINE mode
(/root/reference/train.py:266-289 with train.py:113-115,119-138):

  source step : net.train(); losses, _ = net(image, mask); optim.zero_grad(); loss_ce.mean().backward(); optim.step()
                -- every SyncBatchNorm is in training mode (models/__init__.py:29 -> freeze_bn=False): batch statistics,
                running statistics updated with momentum 0.1
  target pass : with torch.no_grad(): net(image_t, mask_t)  -- only the BN running statistics change ("adaptive BN")
  eval forward: net.eval(); net(image)  -- logits with the adapted running statistics

Weights / inputs come from da_sac_b200.synth (seeded; regenerated identically in the tests).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from da_sac_b200 import synth  # noqa: E402

N_SRC, N_TGT, HW = 4, 3, (129, 129)
ARCH = sys.argv[1] if len(sys.argv) > 1 else "resnet101"          # resnet101 | vgg16 | fcn
if ARCH != "resnet101":
    N_SRC, N_TGT, HW = 3, 2, (96, 96)


def build_reference_net():
    sys.path.insert(0, REF)
    from core.config import cfg, cfg_from_file, cfg_from_list
    cfg_from_file(os.path.join(REF, {"resnet101": "configs/deeplabv2_resnet101_train.yaml", "vgg16": "configs/deeplabv2_vgg16_train.yaml",
                                     "fcn": "configs/fcn_vgg16_train.yaml"}[ARCH]))
    cfg_from_list(["MODEL.INIT_MODEL", "", "MODEL.BASELINE", "True"])
    from models import get_model
    extra = {"drop_rate": 0.0} if ARCH == "fcn" else {}          # Dropout2d off: the comparison must not depend on the RNG stream
    net = get_model(cfg.MODEL, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"), **extra)
    sys.path.remove(REF)
    return net, cfg


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    net, cfg = build_reference_net()
    assert type(net).__name__ == "SAC_Baseline"
    sd0 = {"resnet101": lambda: synth.make_backbone_params(seed=123), "vgg16": lambda: synth.make_vgg16_params(seed=321),
           "fcn": lambda: synth.make_fcn_params(seed=213)}[ARCH]()
    net.backbone.load_state_dict(sd0, strict=True)
    net.train()
    bns = [m for m in net.backbone.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    assert bns and all(m.training for m in bns), "BN layers must train in BASELINE mode"
    optim = torch.optim.SGD(net.parameter_groups(cfg.MODEL.LR, cfg.MODEL.WEIGHT_DECAY), momentum=cfg.MODEL.MOMENTUM)
    # VGG16_FCN8s.forward returns logits_up only when labels are given (fcn.py:139-149): record what _backbone returned
    # (a second call would move the BN running statistics a second time)
    captured = []
    if ARCH == "fcn":
        orig = net.backbone._backbone
        net.backbone._backbone = lambda inp: (captured.append(orig(inp)) or captured[-1])

    def logits_of(o):
        return (o["logits"] if "logits" in o else captured[-1]).detach().numpy()
    xs, ys = synth.make_source_batch(N_SRC, HW, seed=0)
    xt, yt = synth.make_source_batch(N_TGT, HW, seed=1)
    out = {}

    # ---- source step (train.py:119-138)
    losses, outs = net(xs.clone(), ys.clone())
    optim.zero_grad()
    losses["loss_ce"].mean().backward()
    out["src_logits"] = logits_of(outs)
    out["src_loss_ce"] = losses["loss_ce"].detach().numpy()
    names, norms = [], []
    for k, p in net.backbone.named_parameters():
        names.append(k); norms.append(p.grad.double().norm().item())
    out["grad_names"] = np.array(names)
    out["src_grad_norms"] = np.array(norms)
    GRAD_KEYS = {"resnet101": ("model.conv1.weight", "model.bn1.weight", "model.bn1.bias", "model.layer1.0.conv1.weight", "model.layer1.0.bn3.weight",
                               "model.layer2.0.downsample.0.weight", "model.layer2.0.downsample.1.bias", "model.layer3.5.conv2.weight",
                               "model.layer3.5.bn2.weight", "model.layer3.22.bn3.bias", "model.layer4.2.conv3.weight",
                               "model.layer5.conv2d_list.1.weight", "model.layer5.conv2d_list.1.bias"),
                 "vgg16": ("features.0.weight", "features.0.bias", "features.1.weight", "features.1.bias", "features.17.weight", "features.18.bias",
                           "features.39.weight", "features.40.weight", "features.42.weight", "features.42.bias", "features.44.bias",
                           "classifier.conv2d_list.1.weight", "classifier.conv2d_list.1.bias"),
                 "fcn": ("block1.0.weight", "block1.1.weight", "block1.1.bias", "block2.27.weight", "block2.27.bias", "block3.40.weight",
                         "block3.41.bias", "vgg_head.0.weight", "vgg_head.1.weight", "vgg_head.4.bias", "vgg_head.5.bias", "vgg_head.8.weight",
                         "vgg_head.8.bias", "score_pool4.weight", "score_pool3.bias")}[ARCH]
    POST_KEYS = {"resnet101": ("model.layer3.5.conv2.weight", "model.layer3.5.bn2.weight", "model.layer5.conv2d_list.1.bias"),
                 "vgg16": ("features.17.weight", "features.18.weight", "classifier.conv2d_list.1.bias"),
                 "fcn": ("block2.27.weight", "vgg_head.1.weight", "score_pool4.bias")}[ARCH]
    NBT_KEY = {"resnet101": "model.layer3.5.bn2.num_batches_tracked", "vgg16": "features.18.num_batches_tracked",
               "fcn": "vgg_head.1.num_batches_tracked"}[ARCH]
    for k in GRAD_KEYS:
        g = dict(net.backbone.named_parameters())[k].grad
        out["src_grad::" + k] = (g.flatten()[:60000] if g.numel() > 60000 else g).numpy().copy()
    optim.step()
    sd = net.backbone.state_dict()
    stat_keys = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
    out["stat_names"] = np.array(stat_keys)
    out["src_stats"] = np.concatenate([sd[k].numpy().ravel() for k in stat_keys])
    out["src_nbt"] = np.array(int(sd[NBT_KEY]))
    for k in POST_KEYS:
        out["src_post::" + k] = sd[k].flatten()[:60000].numpy().copy()

    # ---- ABN target pass (train.py:281-289)
    with torch.no_grad():
        losses_t, outs_t = net(xt.clone(), yt.clone())
    out["tgt_logits"] = logits_of(outs_t)
    out["tgt_loss_ce"] = losses_t["loss_ce"].numpy()
    sd = net.backbone.state_dict()
    out["tgt_stats"] = np.concatenate([sd[k].numpy().ravel() for k in stat_keys])
    out["tgt_nbt"] = np.array(int(sd[NBT_KEY]))

    # ---- evaluation with the adapted statistics
    net.eval()
    with torch.no_grad():
        logits_e, _ = net.backbone(xt.clone())
    out["eval_logits"] = logits_e.numpy()
    print("src loss %.6f  tgt loss %.6f  logits absmax %.2f / %.2f / %.2f" % (
        float(out["src_loss_ce"]), float(out["tgt_loss_ce"]), np.abs(out["src_logits"]).max(), np.abs(out["tgt_logits"]).max(),
        np.abs(out["eval_logits"]).max()))
    path = os.path.join(HERE, {"resnet101": "abn_resnet101_tiny.npz", "vgg16": "abn_vgg16_tiny.npz", "fcn": "abn_fcn8s_tiny.npz"}[ARCH])
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
