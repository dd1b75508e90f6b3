"""Golden vectors for the FRACTIONAL-GROUP path, produced by the REAL reference on two gloo ranks (CPU).

    python tests/golden/make_golden_fractional.py          # build container only (/root/reference needed)

The reference's default recipe (configs/deeplabv2_resnet101_train.yaml: NUM_GROUPS 2 x GROUP_SIZE 4 on 4 GPUs)
gives every GPU only HALF of a view-group: ``Trainer._prep_batch`` (train.py:185-209) all-gathers the loader
batches and slices, ``SAC._avg_pool`` -> ``_gather`` (models/sac.py:198-216,243-245) all-gathers the
reference-frame probabilities of the ranks that share a group.  Here: 1 group x K=4 views of 96x96 on world_size 2,
rank r holds views [2r, 2r+2).  Writes tests/golden/sac_fractional_w2.npz with, per rank: the teacher logits,
refined probabilities (sub-sampled), confidences, pseudo labels, running_conf, losses.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
K, HW, WORLD = 4, (96, 96), 2


def worker(rank, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.manual_seed(0)
    torch.set_num_threads(4)
    from da_sac_b200 import synth
    sys.path.insert(0, REF)
    from core.config import cfg, cfg_from_file, cfg_from_list
    cfg_from_file(os.path.join(REF, "configs/deeplabv2_resnet101_train.yaml"))
    cfg_from_list(["TRAIN.GROUP_SIZE", str(K), "TRAIN.NUM_GROUPS", "1", "DATASET.CROP_SIZE", "(%d,%d)" % HW,
                   "MODEL.INIT_MODEL", ""])
    from models import get_model
    net = get_model(cfg.MODEL, rank, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    sys.path.remove(REF)
    assert net.world_size == WORLD
    net.backbone.load_state_dict(synth.make_backbone_params(seed=123), strict=True)
    net.train()
    batch = synth.make_target_batch(1, K, HW, seed=3)
    per = K // WORLD
    x, y, x2, A, Ai = [t[rank * per:(rank + 1) * per].clone() for t in batch]       # what _prep_batch hands this rank
    losses, outs = net(x, y, x2, A, Ai, use_teacher=True, update_teacher=True, T=K)
    with torch.no_grad():
        tl, _ = net.slow_net(x2)
    res = {"teacher_logits": tl.numpy(), "teacher_refined_sub": outs["teacher_refined"][:, :, ::3, ::3].contiguous().numpy(),
           "teacher_conf": outs["teacher_conf"].numpy(), "teacher_labels": outs["teacher_labels"].numpy().astype(np.uint8),
           "running_conf": outs["running_conf"].clone().numpy(), "self_ce": losses["self_ce"].detach().numpy(),
           "logits": outs["logits"].detach().numpy()}
    refined = outs["teacher_refined"]
    conf, idx = refined.max(1)
    top2 = refined.topk(2, dim=1).values
    B, C = refined.shape[:2]
    peaks = torch.zeros_like(refined).scatter_(1, idx[:, None], conf[:, None]).view(B, C, -1).max(-1).values
    thr = (peaks * cfg.MODEL.RUN_CONF_UPPER * (1 - torch.exp(-outs["running_conf"] / cfg.MODEL.THRESHOLD_BETA)).view(1, C)).clamp(cfg.MODEL.RUN_CONF_LOWER)
    thr_px = thr.gather(1, idx.view(B, -1)).view_as(conf)
    res["ambiguous"] = (((conf - thr_px).abs() < 1e-5) | (((top2[:, 0] - top2[:, 1]) < 1e-5) & (conf > 0))).numpy()
    lab = outs["teacher_labels"]
    print("rank", rank, "self_ce %.6f" % float(losses["self_ce"]), "valid %.3f" % (lab != 255).float().mean().item(),
          "classes", len(torch.unique(lab[lab != 255])), "ambiguous", int(res["ambiguous"].sum()), flush=True)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs: p.start()
    got = dict(q.get(timeout=1800) for _ in procs)
    for p in procs: p.join(timeout=120)
    out = {}
    for r, res in got.items():
        for k, v in res.items():
            out["r%d_%s" % (r, k)] = v
    path = os.path.join(HERE, "sac_fractional_w2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
