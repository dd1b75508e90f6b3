"""The two batched per-step layout kernels, directly (they are otherwise only seen through whole training steps):

  * ``sacb_prepare_batched``: OIHW fp32 weights -> BN fold (scale / shift vectors) + fprop planes ``[RS][Kf][C]`` (optionally
    carrying the folded scale) + dgrad planes ``[RS][C][Kt]`` (taps flipped, scale folded) as bf16 hi/lo pairs; what
    ``nn.Conv2d`` + eval-mode BN hold implicitly (/root/reference/models/deeplabv2.py:59-99, basenet.py:86-139);
  * ``sacb_wgrad_finalize_batched``: split-K partial planes ``[split][K][RS][C]`` -> ``dW`` in OIHW (sum in split order, times
    gamma/sigma), d(gamma), d(beta), d(bias)  (DESIGN 4.2).
Shapes cover both code paths of each kernel (tiled / element-wise prepare: up to 9 taps / 49 taps; staged / direct finalize:
3x3 up to 512 channels / everything else), ragged tiles (K, C not multiples of 32; padded Kf / Kt) and items without BN.
Planes and dW must be BIT-exact against a torch restatement; d(gamma) to 1e-5 (a dot product whose order is the kernel's)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
EPS = 1e-5

#        K    C   R  Kf   Kt   bn     bias   fold_wf  with_wt
PREP = [(64, 64, 3, 64, 64, True, False, 1, True),
        (256, 64, 1, 256, 256, True, False, 1, True),
        (19, 96, 3, 32, 64, False, True, 0, True),          # ragged: K < Kf < Kt, C = 3 tiles of 32
        (40, 48, 1, 64, 64, True, True, 0, True),           # C not a multiple of 32, fold only in the dgrad planes
        (128, 128, 3, 128, 128, True, False, 1, False),     # teacher: no dgrad planes
        (24, 32, 7, 32, 64, True, False, 1, True),          # 49 taps: the element-wise path
        (64, 64, 2, 64, 64, True, False, 1, True)]          # 4 taps (even tap count: bank-conflicting but correct)


def split(v):
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16)
    return hi.view(torch.int16), lo.view(torch.int16)


def test_prepare_batched_planes_bit_exact():
    from da_sac_b200 import lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(0)
    dev = "cuda"
    items, blocks, keep, expect = [], [], [], []
    for K, Cc, R, Kf, Kt, bn, bias, fold, with_wt in PREP:
        RS = R * R
        w = torch.randn(K, Cc, R, R, generator=g)
        gamma, beta = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g)
        mean, var = torch.randn(K, generator=g), torch.rand(K, generator=g) + 0.1
        b = torch.randn(K, generator=g)
        t = dict(w=w.to(dev), gamma=gamma.to(dev), beta=beta.to(dev), mean=mean.to(dev), var=var.to(dev), b=b.to(dev),
                 scale=torch.full((K,), -7.0, device=dev), shift=torch.full((K,), -7.0, device=dev),
                 fh=torch.full((RS, Kf, Cc), 0x7FC0, dtype=torch.int16, device=dev), fl=torch.full((RS, Kf, Cc), 0x7FC0, dtype=torch.int16, device=dev),
                 th=torch.full((RS, Cc, Kt), 0x7FC0, dtype=torch.int16, device=dev), tl=torch.full((RS, Cc, Kt), 0x7FC0, dtype=torch.int16, device=dev))
        keep.append(t)
        items.append(L.PrepItem(L.dptr(t["w"]), L.dptr(t["gamma"]) if bn else None, L.dptr(t["beta"]) if bn else None,
                                L.dptr(t["mean"]) if bn else None, L.dptr(t["var"]) if bn else None, L.dptr(t["b"]) if bias else None,
                                L.dptr(t["scale"]), L.dptr(t["shift"]), L.dptr(t["fh"]), L.dptr(t["fl"]),
                                L.dptr(t["th"]) if with_wt else None, L.dptr(t["tl"]) if with_wt else None,
                                K, Cc, R, R, Kf, Kt, fold, 0))
        blocks.append(lib.sacb_prep_item_blocks(K, Cc, R, R, Kf, Kt, 1, 1 if with_wt else 0))
        # ---- torch restatement (fp32, the kernel's expressions)
        bias_v = b if bias else torch.zeros(K)
        if bn:
            sc = gamma * (1.0 / torch.sqrt(var + EPS))
            sh = beta + (bias_v - mean) * sc
        else:
            sc, sh = torch.ones(K), bias_v
        expect.append((sc, sh, w))
    tab, begin, n, total = L.item_table(items, blocks, torch.device(dev))
    n0 = L.launch_count()
    L.check(lib.sacb_prepare_batched(L.ptr(tab), L.ptr(begin), n, total, C.c_float(EPS), L.stream()), "sacb_prepare_batched")
    torch.cuda.synchronize()
    assert L.launch_count() == n0 + 1
    for (K, Cc, R, Kf, Kt, bn, bias, fold, with_wt), t, (sc, sh, w) in zip(PREP, keep, expect):
        tag = "K%d C%d R%d" % (K, Cc, R)
        RS = R * R
        got_sc = t["scale"].cpu()
        assert torch.allclose(got_sc, sc, rtol=3e-7, atol=0), tag                    # sqrt / divide: an ulp between libraries
        assert torch.allclose(t["shift"].cpu(), sh, rtol=1e-6, atol=1e-6), tag       # one FMA contraction of freedom
        # the planes must carry exactly the scale vector the epilogues / the finalize kernel use: bit-exact given it
        wk = w.reshape(K, Cc, RS)
        wf = torch.zeros(RS, Kf, Cc)
        wf[:, :K] = (wk * got_sc.view(K, 1, 1) if (fold and bn) else wk).permute(2, 0, 1)
        wt = torch.zeros(RS, Cc, Kt)
        wt[:, :, :K] = (wk * got_sc.view(K, 1, 1) if bn else wk).flip(2).permute(2, 1, 0)
        wf, wt = split(wf), split(wt)
        assert torch.equal(t["fh"].cpu(), wf[0]) and torch.equal(t["fl"].cpu(), wf[1]), tag + " fprop planes"
        if with_wt:
            assert torch.equal(t["th"].cpu(), wt[0]) and torch.equal(t["tl"].cpu(), wt[1]), tag + " dgrad planes"
        else:
            assert bool((t["th"] == 0x7FC0).all()), tag + ": dgrad planes of an item without them were touched"


#        K    C    RS  splits  bn    bias
FIN = [(64, 64, 9, 5, True, False),        # staged transpose
       (96, 512, 9, 3, True, False),       # the largest staged shape
       (256, 64, 1, 7, True, False),       # 1x1: 16-byte stores
       (19, 2048, 9, 2, False, True),      # too large to stage: direct 16-byte path, no BN
       (32, 24, 49, 4, True, True),        # 49 taps
       (16, 30, 9, 3, True, False)]        # C % 4 != 0: scalar path


def test_wgrad_finalize_batched_bit_exact_dw():
    from da_sac_b200 import lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(1)
    dev = "cuda"
    items, blocks, keep, expect = [], [], [], []
    for K, Cc, RS, S, bn, bias in FIN:
        part = torch.randn(S, K, RS, Cc, generator=g)
        w = torch.randn(K, Cc, RS, generator=g)
        scale = torch.rand(K, generator=g) + 0.5
        mean, var, dbeta, b = torch.randn(K, generator=g), torch.rand(K, generator=g) + 0.1, torch.randn(K, generator=g), torch.randn(K, generator=g)
        t = dict(part=part.to(dev), w=w.to(dev), scale=scale.to(dev), mean=mean.to(dev), var=var.to(dev), dbeta=dbeta.to(dev), b=b.to(dev),
                 dw=torch.full((K, Cc, RS), float("nan"), device=dev), dgamma=torch.full((K,), float("nan"), device=dev),
                 dbias=torch.full((K,), float("nan"), device=dev), dbeta_out=torch.full((K,), float("nan"), device=dev))
        keep.append(t)
        items.append(L.FinalizeItem(L.dptr(t["part"]), L.dptr(t["w"]), L.dptr(t["scale"]) if bn else None,
                                    L.dptr(t["mean"]) if bn else None, L.dptr(t["var"]) if bn else None, L.dptr(t["dbeta"]),
                                    L.dptr(t["dw"]), L.dptr(t["dgamma"]) if bn else None, L.dptr(t["b"]) if bias else None,
                                    L.dptr(t["dbias"]) if bias else None, L.dptr(t["dbeta_out"]) if bn else None, K, Cc, RS, S))
        blocks.append(K)
        gs = torch.zeros(K, RS, Cc)
        for sp in range(S):
            gs = gs + part[sp]                                         # split order, fp32
        sc = scale if bn else torch.ones(K)
        dw = (sc.view(K, 1, 1) * gs).permute(0, 2, 1).contiguous()     # [k][c][rs]
        dot = (w.double() * gs.permute(0, 2, 1).double()).sum(dim=(1, 2))
        bb = b.double() if bias else torch.zeros(K, dtype=torch.double)
        dgamma = (dot + (bb - mean.double()) * dbeta.double()) / torch.sqrt(var.double() + EPS)
        mag = ((w.double() * gs.permute(0, 2, 1).double()).abs().sum(dim=(1, 2)) + ((bb - mean.double()) * dbeta.double()).abs()) / torch.sqrt(var.double() + EPS)
        expect.append((dw, dgamma, mag, sc * dbeta))
    tab, begin, n, total = L.item_table(items, blocks, torch.device(dev))
    L.check(lib.sacb_wgrad_finalize_batched(L.ptr(tab), L.ptr(begin), n, total, C.c_float(EPS), L.stream()), "sacb_wgrad_finalize_batched")
    torch.cuda.synchronize()
    for (K, Cc, RS, S, bn, bias), t, (dw, dgamma, mag, dbias) in zip(FIN, keep, expect):
        tag = "K%d C%d RS%d" % (K, Cc, RS)
        assert torch.equal(t["dw"].cpu(), dw), tag + ": dW must be the split-ordered sum, bit for bit"
        if bn:
            err = ((t["dgamma"].cpu().double() - dgamma).abs() / mag).max().item()
            assert err < 1e-5, (tag, err)
            assert torch.equal(t["dbeta_out"].cpu(), t["dbeta"].cpu()), tag
        if bias:
            assert torch.equal(t["dbias"].cpu(), dbias), tag
