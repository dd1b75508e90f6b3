"""Tail split of the CTA-pair GEMM kernel (SACB_TAIL_SPLIT=1; conv_gemm_pair_kernel<.., .., true>): the tiles of the last partial
wave run as two 256 x 128 halves (MMA N = 128).  Same k order per output element, so the planes must be bit-identical to the
default kernel's; also checked against fp64.  The switch is read once per process -> subprocesses.

Green on a B200 since round 2 (profiles/r2a_test_tail_split_gpu.log); the switch stays off by default because it measured slower."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r'''
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, %r)
from da_sac_b200 import lib as L

def split(x):
    hi = x.to(torch.bfloat16); lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()

def nhwc(x): return x.permute(0, 2, 3, 1).contiguous()

out = []
# 6 x 65 x 65 = 25350 rows = 100 pair tiles per 256 output channels: 74 + 26 on a 148-SM part -> the last 26 are split
for (N, H, W, C, K, R, dil, seed) in [(6, 65, 65, 256, 256, 3, 2, 1), (6, 65, 65, 1024, 256, 1, 1, 2), (6, 65, 65, 128, 256, 1, 1, 3)]:
    torch.manual_seed(seed)
    pad = dil * (R // 2)
    x = torch.randn(N, C, H, W, device="cuda"); w = torch.randn(K, C, R, R, device="cuda") / (C * R * R) ** 0.5
    scale = torch.rand(K, device="cuda") + 0.5; shift = torch.randn(K, device="cuda") * 0.1
    ref = F.relu(F.conv2d(x.double(), w.double(), None, 1, pad, dil) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1))
    xh, xl = split(nhwc(x)); wt = w.permute(2, 3, 0, 1).reshape(R * R, K, C).contiguous(); wh, wl = split(wt)
    oh = torch.empty(N, H, W, K, device="cuda", dtype=torch.bfloat16); ol = torch.empty_like(oh)
    cs = torch.zeros(K, device="cuda")
    L.conv_gemm(xh, xl, wh, wl, (N, H, W, C, K, R, 1, dil, pad), scale=scale, shift=shift, relu=True, out_hi=oh, out_lo=ol, colsum=cs)
    torch.cuda.synchronize()
    got = (oh.float() + ol.float()).permute(0, 3, 1, 2).double()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    cerr = ((cs.double() - ref.sum((0, 2, 3))).abs().max() / ref.sum((0, 2, 3)).abs().max()).item()
    print("tail_split=%%s C%%d K%%d %%dx%%d: err %%.2e colsum err %%.2e" %% (sys.argv[1], C, K, R, R, err, cerr))
    assert err < 2e-5 and cerr < 1e-4, (err, cerr)
    out.append((oh.cpu(), ol.cpu()))
torch.save(out, sys.argv[2])
''' % ROOT


def test_tail_split_matches_fp64_and_the_default_kernel(tmp_path):
    import torch
    outs = {}
    for flag in ("0", "1"):
        path = str(tmp_path / ("planes%s.pt" % flag))
        r = subprocess.run([sys.executable, "-c", CHECK, flag, path], env=dict(os.environ, SACB_TAIL_SPLIT=flag),
                           capture_output=True, text=True, timeout=600)
        print(r.stdout, r.stderr[-2000:])
        assert r.returncode == 0, r.stderr[-2000:]
        outs[flag] = torch.load(path)
    for (h0, l0), (h1, l1) in zip(outs["0"], outs["1"]):
        assert torch.equal(h0, h1) and torch.equal(l0, l1), "tail-split and default kernels must produce identical planes"
