"""Golden vectors for the target-view augmentation, produced by the REAL reference transforms (PIL / torchvision).

    python tests/golden/make_golden_aug.py          # build container only (/root/reference needed)

For every case: ``random.seed(s); torch.manual_seed(s)``, then exactly what ``DataTarget.__getitem__`` does after the base
crop exists (dataloader_target.py:281-306): tf_pre tail = GuidedRandHFlip + MaskRandScaleCrop, deepcopy, tf_augm =
RandGaussianBlur + MaskRandJitter + MaskRandGreyscale on copy #1, tf_post = ToTensorMask + Normalize + ApplyMask(-1) on
both, ``_get_affine`` / ``_get_affine_inv``.  Stored next to the inputs so that tests can check (a) that
da_sac_b200.augment.draw_group_params consumes the generators identically (affine operators bit-equal), (b) masks / labels
exactly, (c) pixels to about one grey level.  Writes tests/golden/aug_reference.npz.
"""
import copy
import os
import random
import sys
import types

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

CASES = [dict(seed=1, K=4, hw=(96, 128), zoom=(0.5, 1.0)), dict(seed=2, K=3, hw=(128, 128), zoom=(0.5, 1.0)),
         dict(seed=5, K=4, hw=(80, 112), zoom=(0.5, 1.2)), dict(seed=8, K=2, hw=(64, 64), zoom=(0.5, 1.0))]


def make_base(hw, seed):
    g = np.random.RandomState(seed)
    H, W = hw
    low = g.rand(H // 8 + 2, W // 8 + 2, 3)
    img = np.asarray(Image.fromarray((low * 255).astype(np.uint8)).resize((W, H), Image.BICUBIC)).astype(np.float32)
    img = np.clip(img + g.randn(H, W, 3) * 6, 0, 255).astype(np.uint8)
    mask = np.zeros((H, W), np.uint8)
    mask[:, W - 9:] = 1                                   # MaskRandCrop padding strip
    mask[:5, :] = 1
    img[mask > 0] = 0
    label = (g.randint(0, 19, (H // 16 + 1, W // 16 + 1)).repeat(16, 0).repeat(16, 1)[:H, :W]).astype(np.uint8)
    label[g.rand(H, W) < 0.05] = 255
    return img, mask, label


def main():
    sys.path.insert(0, REF)
    import datasets.tf_target as tf
    from datasets.dataloader_target import DataTarget
    sys.path.remove(REF)
    # NumPy 2 shim for tf_target.py:36 (np.array(pic, np.int32, copy=False))
    tf.ToTensorMask._ToTensorMask__toByteTensor = lambda self, pic: torch.from_numpy(np.asarray(pic, np.int32).copy())
    from da_sac_b200 import augment as AUG
    out = {"n_cases": np.array(len(CASES))}
    for ci, case in enumerate(CASES):
        K, hw, seed = case["K"], case["hw"], case["seed"]
        img, mask, label = make_base(hw, 100 + seed)
        pre = tf.Compose([tf.GuidedRandHFlip(), tf.MaskRandScaleCrop(case["zoom"])])
        augm = tf.Compose([tf.RandGaussianBlur(), tf.MaskRandJitter(0.4), tf.MaskRandGreyscale(0.2)])
        post = tf.Compose([tf.ToTensorMask(), tf.Normalize(mean=AUG.MEAN, std=AUG.STD), tf.ApplyMask(-1)])
        random.seed(seed); torch.manual_seed(seed)
        images = [Image.fromarray(img).copy() for _ in range(K)]
        labels = [Image.fromarray(label).copy() for _ in range(K)]
        masks = [Image.fromarray(mask).copy() for _ in range(K)]
        augms = pre(images, labels, masks)
        affine_params = augms[-1]
        augms = augms[:-1]
        augms2 = copy.deepcopy(augms)
        augms1 = augm(*augms)
        images1, gts = post(*augms1)
        images2, _ = post(*augms2)
        stub = types.SimpleNamespace(cfg=types.SimpleNamespace(TRAIN=types.SimpleNamespace(GROUP_SIZE=K),
                                                               DATASET=types.SimpleNamespace(CROP_SIZE=list(hw))))
        affine = DataTarget._get_affine(stub, affine_params)
        affine_inv = DataTarget._get_affine_inv(stub, affine, affine_params)
        # our draw, same seeds
        cfg = type("Cfg", (AUG.AugCfg,), {"RND_ZOOM": case["zoom"]})
        random.seed(seed); torch.manual_seed(seed)
        rows, aff = AUG.draw_group_params(K, hw, cfg)
        pre_ = "c%d_" % ci
        out.update({pre_ + "base": img, pre_ + "base_mask": mask, pre_ + "base_label": label,
                    pre_ + "rows": np.asarray(rows, np.float32), pre_ + "zoom": np.asarray(case["zoom"], np.float64),
                    pre_ + "seed": np.array(seed), pre_ + "K": np.array(K),
                    pre_ + "frames1": torch.stack(images1).numpy(), pre_ + "frames2": torch.stack(images2).numpy(),
                    pre_ + "gt": torch.stack(gts).numpy().astype(np.int16),
                    pre_ + "affine": affine.numpy(), pre_ + "affine_inv": affine_inv.numpy(),
                    pre_ + "affine_params": np.asarray(affine_params, np.float64)})
        print("case", ci, case, "affine params", [[round(v, 3) for v in p] for p in affine_params])
        print("   ours         ", [[round(v, 3) for v in p] for p in aff])
    path = os.path.join(HERE, "aug_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
