"""BaseNet: optimiser-group / BN-freezing bookkeeping of the backbones.

Mirrors the public surface of /root/reference/models/basenet.py:12-143
(``from_scratch_layers``, ``bn_freeze``, ``train()`` keeping frozen BN in eval,
``parameter_groups(base_lr, wd)`` -> 4 groups with lr multipliers), because
``train.py:93-94`` and ``base_trainer.get_optim`` consume exactly that."""
import torch.nn as nn


class BaseNet(nn.Module):
    _trainable = (nn.Linear, nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d, nn.GroupNorm, nn.InstanceNorm2d, nn.SyncBatchNorm)
    _batchnorm = (nn.BatchNorm2d, nn.SyncBatchNorm, nn.GroupNorm)

    def __init__(self):
        super().__init__()
        self.from_scratch_layers = []
        self.not_training = []
        self.bn_freeze = []

    def lr_mult(self):
        return 1., 1.

    def lr_mult_bias(self):
        return 2., 2.

    def _is_learnable(self, layer):
        return isinstance(layer, BaseNet._trainable)

    def _from_scratch(self, net):
        for layer in net.modules():
            if self._is_learnable(layer):
                self.from_scratch_layers.append(layer)

    def _freeze_bn(self, net):
        for layer in net.modules():
            if isinstance(layer, BaseNet._batchnorm):
                self.bn_freeze.append(layer)

    def train(self, mode=True):
        super().train(mode)
        for layer in self.bn_freeze:      # frozen BN: statistics stay in eval mode (basenet.py:97-100)
            layer.eval()
        return self

    def parameter_groups(self, base_lr, wd):
        """[old weights (wd), old biases, new weights x lr_mult (wd), new biases] -- basenet.py:102-139;
        BN gamma lands in the weight groups, BN beta in the bias groups."""
        w_old, w_new = self.lr_mult()
        b_old, b_new = self.lr_mult_bias()
        groups = ({"params": [], "weight_decay": wd, "lr": w_old * base_lr},
                  {"params": [], "weight_decay": 0.0, "lr": b_old * base_lr},
                  {"params": [], "weight_decay": wd, "lr": w_new * base_lr},
                  {"params": [], "weight_decay": 0.0, "lr": b_new * base_lr})
        scratch = set(id(m) for m in self.from_scratch_layers)
        for m in self.modules():
            if not self._is_learnable(m):
                continue
            new = id(m) in scratch
            if m.weight is not None and m.weight.requires_grad:
                groups[2 if new else 0]["params"].append(m.weight)
            if m.bias is not None and m.bias.requires_grad:
                groups[3 if new else 1]["params"].append(m.bias)
        return groups
