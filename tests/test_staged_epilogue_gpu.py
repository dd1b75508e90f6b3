"""Residual-staging variant of the CTA-pair GEMM kernel (SACB_EPI_STAGED=1; csrc/sacb_gemm.cu conv_gemm_pair_kernel<true>):
fprop + BN affine + residual + ReLU of the 1x1 expand layers and the dgrad-with-skip-gradient form, against fp64, and
bit-identical to the default kernel (same MMA order, same epilogue arithmetic -- only where the residual is read from differs).

The switch is read once per process, so the check runs in a subprocess.  Green on a B200 since round 2
(profiles/r2a_test_staged_epilogue_gpu.log)."""
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

def run(N, H, W, C, K, seed, with_mask):
    torch.manual_seed(seed)
    x = torch.randn(N, C, H, W, device="cuda"); w = torch.randn(K, C, 1, 1, device="cuda") / C ** 0.5
    scale = torch.rand(K, device="cuda") + 0.5; shift = torch.randn(K, device="cuda") * 0.1
    res = torch.randn(N, K, H, W, device="cuda")
    xh, xl = split(nhwc(x)); rh, rl = split(nhwc(res))
    wt = w.permute(2, 3, 0, 1).reshape(1, K, C).contiguous(); wh, wl = split(wt)
    act = F.relu(torch.randn(N, H, W, K, device="cuda")); mh, _ = split(act)
    conv = F.conv2d(x.double(), w.double())
    resd = (rh.float() + rl.float()).double().permute(0, 3, 1, 2)
    if with_mask:      # dgrad form: (acc + skip gradient) masked by the ReLU of the layer input, plus its column sums
        ref = (conv + resd) * (act > 0).permute(0, 3, 1, 2)
        kw = dict(add_hi=rh, add_lo=rl, mask_hi=mh)
    else:              # fprop form: relu(acc * scale + shift + residual)
        ref = F.relu(conv * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1) + resd)
        kw = dict(scale=scale, shift=shift, add_hi=rh, add_lo=rl, relu=True)
    oh = torch.empty(N, H, W, K, device="cuda", dtype=torch.bfloat16); ol = torch.empty_like(oh)
    cs = torch.zeros(K, device="cuda")
    L.conv_gemm(xh, xl, wh, wl, (N, H, W, C, K, 1, 1, 1, 0), out_hi=oh, out_lo=ol, colsum=cs, **kw)
    torch.cuda.synchronize()
    got = (oh.float() + ol.float()).permute(0, 3, 1, 2).double()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    cerr = ((cs.double() - ref.sum((0, 2, 3))).abs().max() / ref.sum((0, 2, 3)).abs().max()).item()
    return err, cerr, oh.clone(), ol.clone()

out = []
for (N, H, W, C, K, seed, m) in [(3, 33, 33, 256, 1024, 1, False), (2, 20, 31, 512, 256, 2, False), (3, 33, 33, 256, 1024, 3, True),
                                 (1, 65, 65, 256, 512, 4, True)]:
    err, cerr, oh, ol = run(N, H, W, C, K, seed, m)
    print("staged=%%s N%%d %%dx%%d C%%d K%%d mask=%%s: err %%.2e colsum err %%.2e" %% (sys.argv[1], N, H, W, C, K, m, err, cerr))
    assert err < 5e-5 and cerr < 1e-4, (err, cerr)
    out.append((oh.cpu(), ol.cpu()))
torch.save(out, sys.argv[2])
''' % ROOT


def test_staged_epilogue_matches_fp64_and_the_default_kernel(tmp_path):
    import torch
    outs = {}
    for flag in ("0", "1"):
        path = str(tmp_path / ("planes%s.pt" % flag))
        env = dict(os.environ, SACB_EPI_STAGED=flag)
        r = subprocess.run([sys.executable, "-c", CHECK, flag, path], env=env, capture_output=True, text=True, timeout=600)
        print(r.stdout, r.stderr[-2000:])
        assert r.returncode == 0, r.stderr[-2000:]
        outs[flag] = torch.load(path)
    for (h0, l0), (h1, l1) in zip(outs["0"], outs["1"]):
        assert torch.equal(h0, h1) and torch.equal(l0, l1), "staged and default epilogues must produce identical planes"
