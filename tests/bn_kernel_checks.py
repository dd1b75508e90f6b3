"""Kernel-level checks of the training-mode BN entry points (csrc/sacb_bn.cu) against torch fp64, written once and run
(a) on a B200 through libsac_b200.so (tests/test_abn_gpu.py) and (b) in the GPU-less container through the host emulation of
the SAME kernel source (tests/cpu_emul, tests/test_emul_bn_cpu.py)."""
import ctypes as C

import torch
import torch.nn.functional as F

EPS, MOM = 1e-5, 0.1


class Api:
    """what the checks need from a library: the ctypes handle, the stream argument, a pointer maker and a synchronise"""

    def __init__(self, lib, stream, ptr, sync, dev):
        self.lib, self.st, self.ptr, self.sync, self.dev = lib, stream, ptr, sync, dev

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.sacb_last_error().decode()))


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def join(hi, lo):
    return hi.float() + lo.float()


def rel(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


CASES = [(1000, 64, False), (5003, 256, True), (513, 1024, False)]


def bn_kernels_match_torch_fp64(api, M, Cn, with_res):
    lib, st, L = api.lib, api.st, api
    torch.manual_seed(M + Cn)
    dev = api.dev
    z = torch.randn(M, Cn, device=dev) * (torch.rand(Cn, device=dev) * 2 + 0.2) + torch.randn(Cn, device=dev) * 3
    zh, zl = split(z)
    z = join(zh, zl)                                         # what the kernels see
    gamma = torch.rand(Cn, device=dev) + 0.5; beta = torch.randn(Cn, device=dev)
    rm = torch.randn(Cn, device=dev) * 0.1; rv = torch.rand(Cn, device=dev) + 0.5
    res = torch.randn(M, Cn, device=dev) if with_res else None
    rh, rl = split(res) if with_res else (None, None)
    if with_res:
        res = join(rh, rl)
    # reference in fp64
    zd = z.double().t().reshape(1, Cn, M).requires_grad_(True)
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    y_pre = F.batch_norm(zd, rm_ref, rv_ref, gamma.double(), beta.double(), True, MOM, EPS)
    y_ref = torch.relu(y_pre + (res.double().t().reshape(1, Cn, M) if with_res else 0))
    # kernels: forward
    partials = torch.empty(int(lib.sacb_bn_moments_partial_elems(C.c_int64(M), Cn)), device=dev, dtype=torch.float64)
    sums = torch.empty(2 * Cn, device=dev, dtype=torch.float64)
    mean = torch.empty(Cn, device=dev); invstd = torch.empty(Cn, device=dev); scale = torch.empty(Cn, device=dev)
    L.check(lib.sacb_bn_moments(L.ptr(zh), L.ptr(zl), None, None, None, None, 0, C.c_int64(M), Cn, L.ptr(partials), L.ptr(sums), st),
            "sacb_bn_moments")
    assert rel(sums[:Cn], z.double().sum(0))[1] < 1e-12 and rel(sums[Cn:], (z.double() ** 2).sum(0))[1] < 1e-12   # fp64 end to end
    rm_k, rv_k = rm.clone(), rv.clone()
    L.check(lib.sacb_bn_train_finalize(L.ptr(sums), C.c_double(float(M)), L.ptr(gamma), C.c_float(EPS), C.c_float(MOM), L.ptr(rm_k),
                                       L.ptr(rv_k), L.ptr(mean), L.ptr(invstd), L.ptr(scale), Cn, st), "sacb_bn_train_finalize")
    assert rel(mean, z.double().mean(0))[1] < 1e-6
    assert rel(invstd, 1.0 / (z.double().var(0, unbiased=False) + EPS).sqrt())[1] < 1e-5
    assert rel(rm_k, rm_ref)[1] < 1e-6 and rel(rv_k, rv_ref)[1] < 1e-5
    yh = torch.empty(M, Cn, device=dev, dtype=torch.bfloat16); yl = torch.empty_like(yh)
    L.check(lib.sacb_bn_apply(L.ptr(zh), L.ptr(zl), L.ptr(mean), L.ptr(scale), L.ptr(beta), L.ptr(rh), L.ptr(rl), 1, L.ptr(yh),
                              L.ptr(yl), C.c_int64(M), Cn, st), "sacb_bn_apply")
    assert rel(join(yh, yl), y_ref.detach().reshape(Cn, M).t())[1] < 2e-5
    # backward: g at the BN output (after the ReLU mask) -> dz, d gamma, d beta
    g = torch.randn(M, Cn, device=dev) * (join(yh, yl) > 0).float()
    gh, gl = split(g)
    g = join(gh, gl)
    y_pre.backward(g.double().t().reshape(1, Cn, M))
    L.check(lib.sacb_bn_moments(L.ptr(gh), L.ptr(gl), L.ptr(zh), L.ptr(zl), L.ptr(mean), L.ptr(invstd), 1, C.c_int64(M), Cn,
                                L.ptr(partials), L.ptr(sums), st), "sacb_bn_moments(bwd)")
    dgamma = torch.empty(Cn, device=dev); dbeta = torch.empty(Cn, device=dev); coef = torch.empty(3 * Cn, device=dev)
    L.check(lib.sacb_bn_bwd_finalize(L.ptr(sums), L.ptr(sums), C.c_double(float(M)), L.ptr(gamma), L.ptr(invstd), L.ptr(dgamma),
                                     L.ptr(dbeta), L.ptr(coef), Cn, st), "sacb_bn_bwd_finalize")
    dzh = torch.empty_like(gh); dzl = torch.empty_like(gl)
    L.check(lib.sacb_bn_bwd_apply(L.ptr(gh), L.ptr(gl), L.ptr(zh), L.ptr(zl), L.ptr(mean), L.ptr(invstd), L.ptr(coef), L.ptr(dzh),
                                  L.ptr(dzl), C.c_int64(M), Cn, st), "sacb_bn_bwd_apply")
    api.sync()
    assert rel(dbeta, g.double().sum(0))[1] < 1e-6
    xhat = (z.double() - z.double().mean(0)) / (z.double().var(0, unbiased=False) + EPS).sqrt()
    assert rel(dgamma, (g.double() * xhat).sum(0))[1] < 1e-5
    assert rel(join(dzh, dzl), zd.grad.reshape(Cn, M).t())[1] < 5e-5
    # in place (dz aliases g) gives the same bits
    L.check(lib.sacb_bn_bwd_apply(L.ptr(gh), L.ptr(gl), L.ptr(zh), L.ptr(zl), L.ptr(mean), L.ptr(invstd), L.ptr(coef), L.ptr(gh),
                                  L.ptr(gl), C.c_int64(M), Cn, st), "sacb_bn_bwd_apply(in place)")
    assert torch.equal(gh, dzh) and torch.equal(gl, dzl)


def bn_moments_survive_large_mean(api):
    """|mean| = 600 x std: E[z^2] - E[z]^2 loses 5.5 digits; invstd must still match fp64 to 1e-5 (fp32 sums of squares do not)"""
    lib, st, L, dev = api.lib, api.st, api, api.dev
    M, Cn = 3000, 32
    torch.manual_seed(5)
    zh, zl = split(30.0 + 0.05 * torch.randn(M, Cn, device=dev))
    z = join(zh, zl).double()
    partials = torch.empty(int(lib.sacb_bn_moments_partial_elems(C.c_int64(M), Cn)), device=dev, dtype=torch.float64)
    sums = torch.empty(2 * Cn, device=dev, dtype=torch.float64)
    mean = torch.empty(Cn, device=dev); invstd = torch.empty(Cn, device=dev); scale = torch.empty(Cn, device=dev)
    gamma = torch.ones(Cn, device=dev)
    L.check(lib.sacb_bn_moments(L.ptr(zh), L.ptr(zl), None, None, None, None, 0, C.c_int64(M), Cn, L.ptr(partials), L.ptr(sums), st),
            "sacb_bn_moments")
    L.check(lib.sacb_bn_train_finalize(L.ptr(sums), C.c_double(float(M)), L.ptr(gamma), C.c_float(EPS), C.c_float(MOM), None, None,
                                       L.ptr(mean), L.ptr(invstd), L.ptr(scale), Cn, st), "sacb_bn_train_finalize")
    api.sync()
    assert rel(invstd, 1.0 / (z.var(0, unbiased=False) + EPS).sqrt())[1] < 1e-5
    assert torch.equal(scale, invstd)
