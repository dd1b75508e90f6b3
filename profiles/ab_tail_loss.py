"""Bit-identity + timing probe of the teacher tail and the student loss kernels, for A/B runs of two builds of the library:

    SACB_LIB=da_sac_b200/libsac_b200_prev.so python profiles/ab_tail_loss.py > prev.txt
    python profiles/ab_tail_loss.py > new.txt ; diff <(grep sha prev.txt) <(grep sha new.txt)

Per geometry: sha256 of every output tensor (probabilities, labels, confidence, running_conf, the two losses, d logits) and the
CUDA-event time of the tail / loss forward / loss backward calls (L2 flushed between repetitions)."""
import ctypes as C
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from da_sac_b200 import lib as L, synth  # noqa: E402
from da_sac_b200.models import get_model  # noqa: E402

dev = torch.device("cuda")
GEOMS = [(8, 3, (512, 512)), (2, 4, (640, 640)), (1, 6, (1024, 1024)), (2, 2, (97, 131)), (1, 3, (70, 203)), (2, 2, (128, 128))]


def low_res(n):
    """DeepLabv2-ResNet output size: 7x7 s2 p3 conv, 3x3 s2 p1 ceil-mode pool, one stride-2 stage (deeplabv2.py:125-131)"""
    n = (n - 1) // 2 + 1
    n = -(-(n - 1) // 2) + 1
    return (n - 1) // 2 + 1


def sha(t):
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def timed(fn, flush, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    cfg = synth.ModelCfg()
    m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    m.cuda().train()
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    print("library:", L.LIB_PATH)
    for G, K, HW in GEOMS:
        BT = G * K
        H, W = HW
        g = torch.Generator().manual_seed(H * 7 + W)
        x, y, x2, A, Ai = [t.cuda() for t in synth.make_target_batch(G, K, HW, seed=1)]
        h, w = low_res(H), low_res(W)
        t_logits = (torch.randn(BT, 19, h, w, generator=g) * 3).cuda()
        s_logits = (torch.randn(BT, 19, h, w, generator=g) * 3).cuda()
        tag = "%dx%dx%d K=%d (low-res %dx%d)" % (BT, H, W, K, h, w)

        def tail():
            m.running_conf.fill_(0.02)
            return m._tail(t_logits, y, A, Ai, K)
        ws = tail()
        torch.cuda.synchronize()
        for k in ("probs", "labels", "conf", "conf_mean", "thresholds"):
            print("sha %-34s tail.%-12s %s" % (tag, k, sha(ws[k])))
        print("sha %-34s tail.%-12s %s" % (tag, "running_conf", sha(m.running_conf)))
        t_tail = timed(tail, flush)
        losses = torch.zeros(2, device=dev); scratch = torch.zeros(2, dtype=torch.float64, device=dev)
        grows = torch.empty(BT * 19 * H * w, device=dev)
        dl = torch.zeros_like(s_logits)
        yy = y.clone(); yy[yy == -1] = 255
        d = L.Loss(C.sizeof(L.Loss), BT, 19, h, w, H, W, L.ptr(s_logits), L.ptr(yy), L.ptr(ws["labels"]), L.ptr(ws["conf_mean"]),
                   L.ptr(m.running_conf), 3.0, L.ptr(losses), L.ptr(scratch), 5.0, L.ptr(dl), None, L.ptr(grows))
        fwd = lambda: L.check(L.lib().sacb_student_loss_fwd(C.byref(d), L.stream()), "fwd")
        bwd = lambda: L.check(L.lib().sacb_student_loss_bwd(C.byref(d), L.stream()), "bwd")
        fwd(); bwd(); torch.cuda.synchronize()
        print("sha %-34s loss.%-12s %s   (loss_ce %.6f self_ce %.6f)" % (tag, "losses", sha(losses), losses[0].item(), losses[1].item()))
        print("sha %-34s loss.%-12s %s" % (tag, "dlogits", sha(dl)))
        # the source-pass form (labels == NULL: plain CE against y)
        d2 = L.Loss(C.sizeof(L.Loss), BT, 19, h, w, H, W, L.ptr(s_logits), L.ptr(yy), None, L.ptr(ws["conf_mean"]),
                    L.ptr(m.running_conf), 3.0, L.ptr(losses), L.ptr(scratch), 1.0, L.ptr(dl), None, L.ptr(grows))
        L.check(L.lib().sacb_student_loss_bwd(C.byref(d2), L.stream()), "bwd")
        torch.cuda.synchronize()
        print("sha %-34s loss.%-12s %s" % (tag, "dlogits_ce", sha(dl)))
        print("time %-33s tail %.3f ms, loss fwd %.3f ms, loss bwd %.3f ms" % (tag, t_tail, timed(fwd, flush), timed(bwd, flush)))
        del ws, grows, dl, flush
        torch.cuda.empty_cache()
        flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


main()
