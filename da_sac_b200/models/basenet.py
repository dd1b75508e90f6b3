"""BaseNet: what ``train.py`` / ``base_trainer.get_optim`` expect from a backbone besides ``forward``.

Public surface mirrored from /root/reference/models/basenet.py:12-143: the bookkeeping lists ``from_scratch_layers`` /
``not_training`` / ``bn_freeze``, ``train()`` that leaves frozen BN layers in eval mode, and ``parameter_groups(base_lr, wd)``
returning the four optimiser groups (pre-trained weights, pre-trained biases at 2x lr, new weights and new biases at the
backbone's ``lr_mult`` / ``lr_mult_bias`` factors; weight decay on the weight groups only)."""
import torch.nn as nn

LEARNABLE = (nn.Linear, nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d, nn.SyncBatchNorm, nn.GroupNorm, nn.InstanceNorm2d)
NORMS = (nn.BatchNorm2d, nn.SyncBatchNorm, nn.GroupNorm)


class BaseNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.from_scratch_layers, self.not_training, self.bn_freeze = [], [], []

    # learning-rate multipliers (old layers, from-scratch layers); backbones override them
    def lr_mult(self):
        return 1., 1.

    def lr_mult_bias(self):
        return 2., 2.

    def _is_learnable(self, layer):
        return isinstance(layer, LEARNABLE)

    def _from_scratch(self, net):
        self.from_scratch_layers.extend(m for m in net.modules() if self._is_learnable(m))

    def _freeze_bn(self, net):
        self.bn_freeze.extend(m for m in net.modules() if isinstance(m, NORMS))

    def train(self, mode=True):
        super().train(mode)
        for m in self.bn_freeze:          # frozen statistics: these stay in eval mode whatever the net is told
            m.eval()
        return self

    def parameter_groups(self, base_lr, wd):
        """group index = 2 * is_new + is_bias; BN gamma counts as a weight, BN beta as a bias (as in the reference)"""
        mult = (self.lr_mult()[0], self.lr_mult_bias()[0], self.lr_mult()[1], self.lr_mult_bias()[1])
        groups = [{"params": [], "weight_decay": wd if i % 2 == 0 else 0.0, "lr": mult[i] * base_lr} for i in range(4)]
        new_ids = {id(m) for m in self.from_scratch_layers}
        for m in self.modules():
            if not self._is_learnable(m):
                continue
            base = 2 if id(m) in new_ids else 0
            for is_bias, p in enumerate((m.weight, m.bias)):
                if p is not None and p.requires_grad:
                    groups[base + is_bias]["params"].append(p)
        return tuple(groups)
