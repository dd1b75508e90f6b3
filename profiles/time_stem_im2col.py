"""micro-benchmark + self-check of sacb_stem_im2col (ResNet stem 7x7 s2 and VGG first conv 3x3 s1) against torch unfold"""
import os, sys, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from da_sac_b200 import lib as L
dev = torch.device("cuda")
for (N, H, W, R, stride, pad, KP) in ((24, 512, 512, 7, 2, 3, 192), (8, 512, 512, 3, 1, 1, 64), (3, 97, 131, 7, 2, 3, 192)):
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    x = torch.randn(N, 3, H, W, device=dev)
    hi = torch.empty(N * P * Q * KP, device=dev, dtype=torch.bfloat16); lo = torch.empty_like(hi)
    def run():
        L.check(L.lib().sacb_stem_im2col(L.ptr(x), L.ptr(hi), L.ptr(lo), N, H, W, P, Q, R, stride, pad, KP, L.stream()), "im2col")
    for _ in range(3): run()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ref = torch.nn.functional.unfold(x, R, padding=pad, stride=stride).transpose(1, 2).reshape(N * P * Q, 3 * R * R)
    got = (hi.float() + lo.float()).view(N * P * Q, KP)
    err = (got[:, :3 * R * R] - ref).abs().max().item()
    assert err < 1e-4 and got[:, 3 * R * R:].abs().max().item() == 0.0, err
    print("N=%d %dx%d R=%d s=%d: %.3f ms, %.0f GB/s written, max err %.1e" % (N, H, W, R, stride, e0.elapsed_time(e1) / 10,
          2 * hi.numel() * 2 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9, err))
