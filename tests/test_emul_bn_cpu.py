"""The training-mode BN kernels (csrc/sacb_bn_kernels.cuh + the entry points of csrc/sacb_bn.cu) have not run on a B200 yet
(DESIGN 6i).  Until they have, this file executes the SAME source in the GPU-less container: tests/cpu_emul compiles it for the
host (one std::thread per CUDA thread, std::barrier for __syncthreads, launches run block by block with the grid / block
geometry the entry points compute) and the checks are the ones tests/test_abn_gpu.py runs on the GPU
(tests/bn_kernel_checks.py, torch fp64 reference).  The emulation library is test infrastructure: it is built under
tests/cpu_emul/_build, exports `sacb_emul_marker`, and da_sac_b200/lib.py refuses to load a library that does."""
import ctypes as C
import os
import shutil
import subprocess

import pytest
import torch

import bn_kernel_checks as K

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "cpu_emul")


@pytest.fixture(scope="module")
def api():
    import emul_harness as E
    if not E.available():
        pytest.skip("no host toolchain")
    lib = E.emul_lib()                                         # honours SACB_EMUL_SO (AddressSanitizer build)

    def ptr(t):
        if t is None:
            return None
        assert not t.is_cuda and t.is_contiguous()
        return C.c_void_p(t.data_ptr())

    return K.Api(lib, None, ptr, lambda: None, "cpu")


# the GPU cases (M below / above / not a multiple of the 512-row block; 8 / 32 / 128 channel vectors) plus the edges:
# one channel vector (block wider than the tensor), one row, exactly one row block, a residual on a narrow tensor
@pytest.mark.parametrize("M,Cn,with_res", K.CASES + [(7, 8, False), (1, 16, False), (512, 40, True), (1537, 136, False)])
def test_bn_kernels_emulated_match_torch_fp64(api, M, Cn, with_res):
    if M == 1:
        pytest.skip("torch refuses training-mode BN with one value per channel; covered by test_single_row below")
    n0 = api.lib.sacb_launch_count()
    K.bn_kernels_match_torch_fp64(api, M, Cn, with_res)
    assert api.lib.sacb_launch_count() - n0 == 9          # 2+1+1 forward, 2+1+1 backward, 1 in-place repeat


def test_moments_survive_large_mean(api):
    K.bn_moments_survive_large_mean(api)


def test_single_row(api):
    """M = 1: variance 0, invstd = 1/sqrt(eps), the unbiased correction is skipped (count - 1 = 0)"""
    Cn = 16
    z = torch.randn(1, Cn)
    zh, zl = K.split(z)
    z = K.join(zh, zl)
    partials = torch.empty(int(api.lib.sacb_bn_moments_partial_elems(1, Cn)), dtype=torch.float64)
    sums = torch.empty(2 * Cn, dtype=torch.float64)
    api.check(api.lib.sacb_bn_moments(api.ptr(zh), api.ptr(zl), None, None, None, None, 0, C.c_int64(1), Cn, api.ptr(partials),
                                      api.ptr(sums), None), "moments")
    assert torch.equal(sums[:Cn], z[0].double())
    gamma = torch.ones(Cn); mean = torch.empty(Cn); invstd = torch.empty(Cn); scale = torch.empty(Cn)
    rm = torch.zeros(Cn); rv = torch.ones(Cn)
    api.check(api.lib.sacb_bn_train_finalize(api.ptr(sums), C.c_double(1.0), api.ptr(gamma), C.c_float(K.EPS), C.c_float(K.MOM),
                                             api.ptr(rm), api.ptr(rv), api.ptr(mean), api.ptr(invstd), api.ptr(scale), Cn, None),
              "finalize")
    assert torch.equal(mean, z[0])
    assert torch.allclose(invstd, torch.full((Cn,), K.EPS ** -0.5), rtol=1e-6)
    assert torch.allclose(rm, K.MOM * z[0]) and torch.allclose(rv, torch.full((Cn,), 1 - K.MOM), atol=1e-6)


def test_argument_errors(api):
    """the entry points reject what the kernels cannot index (same messages on the GPU: the checks precede the launch)"""
    t = torch.zeros(64, dtype=torch.float64)
    assert api.lib.sacb_bn_moments(api.ptr(t), api.ptr(t), None, None, None, None, 0, C.c_int64(4), 12, api.ptr(t), api.ptr(t), None) == -1
    assert b"C % 8" in api.lib.sacb_last_error()
    assert api.lib.sacb_bn_moments(api.ptr(t), api.ptr(t), None, None, None, None, 1, C.c_int64(4), 8, api.ptr(t), api.ptr(t), None) == -1
    assert b"mode 1" in api.lib.sacb_last_error()
    assert api.lib.sacb_bn_apply(api.ptr(t), api.ptr(t), api.ptr(t), api.ptr(t), api.ptr(t), api.ptr(t), None, 1, api.ptr(t),
                                 api.ptr(t), C.c_int64(1), 8, None) == -1
    assert b"go together" in api.lib.sacb_last_error()


@pytest.mark.parametrize("name", ["libsacb_emul.so", "libsacb_emul_tc.so", "libsacb_emul_full.so"])
def test_product_loader_refuses_the_emulation_libraries(api, monkeypatch, name):
    """no CPU route into the product: da_sac_b200.lib (hence models/, bench.py, smoke()) rejects anything built from tests/cpu_emul,
    also when SACB_LIB points at it"""
    import emul_harness as E
    from da_sac_b200 import lib as L
    monkeypatch.setattr(L, "LIB_PATH", os.path.join(os.path.dirname(E.SO), name))
    monkeypatch.setattr(L, "_lib", None)
    with pytest.raises(L.SacbError, match="emulation"):
        L.lib()
    assert L._lib is None
