"""ctypes binding of libsac_b200.so (the C ABI declared in include/sacb.h).

The product path has NO fallback: if the shared library is missing or a call
fails, an exception is raised (``SacbError``).  PyTorch is used only for device
memory and streams; the structs below carry raw device pointers.

There is no CPU route either: the host-emulation libraries of tests/cpu_emul
(test infrastructure for the GPU-less build container) export ``sacb_emul_marker``
and ``lib()`` refuses to load anything that does -- also through ``SACB_LIB`` --
and ``ptr()`` rejects tensors that are not on a CUDA device.  Only the tests swap
``lib`` / ``ptr`` for the emulation (tests/emul_harness.py).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SACB_LIB") or os.path.join(_HERE, "libsac_b200.so")      # SACB_LIB: A/B runs of two builds
ABI_VERSION = 6          # SACB_ABI_VERSION in include/sacb.h


class SacbError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SacbError("libsac_b200.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`): " + LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        if hasattr(_lib, "sacb_emul_marker"):
            _lib = None
            raise SacbError("refusing to load a host emulation library (tests/cpu_emul) as libsac_b200: " + LIB_PATH)
        _lib.sacb_last_error.restype = C.c_char_p
        _lib.sacb_launch_count.restype = C.c_int64
        for f in ("sacb_tail_part_sums_elems", "sacb_tail_probs_elems", "sacb_tail_pooled_elems"):
            getattr(_lib, f).restype = C.c_size_t
        if hasattr(_lib, "sacb_bn_moments_partial_elems"):
            _lib.sacb_bn_moments_partial_elems.restype = C.c_size_t
            _lib.sacb_bn_moments_partial_elems.argtypes = [C.c_int64, C.c_int]
        if _lib.sacb_abi_version() != ABI_VERSION:
            raise SacbError("libsac_b200.so ABI version mismatch")
    return _lib


def check(rc, what):
    if rc != 0:
        raise SacbError("%s failed (%d): %s" % (what, rc, lib().sacb_last_error().decode()))


def ptr(t):
    """device pointer of a tensor (or None)"""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libsac_b200 needs contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device(t):
    """True for a CUDA tensor (the tests' host emulation swaps this together with ptr / stream)"""
    return t.is_cuda


def launch_count():
    return int(lib().sacb_launch_count())


_vp = C.c_void_p


class ConvGemm(C.Structure):
    _fields_ = [("size", C.c_uint32),
                ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("K", C.c_int32), ("k_valid", C.c_int32),
                ("R", C.c_int32), ("S", C.c_int32), ("stride", C.c_int32), ("dil", C.c_int32), ("pad", C.c_int32),
                ("P", C.c_int32), ("Q", C.c_int32),
                ("x_hi", _vp), ("x_lo", _vp), ("wt_hi", _vp), ("wt_lo", _vp),
                ("scale", _vp), ("shift", _vp), ("add_f32", _vp), ("add_hi", _vp), ("add_lo", _vp), ("mask_hi", _vp),
                ("relu", C.c_int32),
                ("out_hi", _vp), ("out_lo", _vp), ("out_f32", _vp), ("out_nchw", _vp), ("colsum", _vp),
                ("precision", C.c_int32), ("unit_scale", C.c_int32)]


class ConvWgrad(C.Structure):
    _fields_ = [("size", C.c_uint32),
                ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("K", C.c_int32), ("k_valid", C.c_int32),
                ("R", C.c_int32), ("S", C.c_int32), ("stride", C.c_int32), ("dil", C.c_int32), ("pad", C.c_int32),
                ("P", C.c_int32), ("Q", C.c_int32),
                ("x_hi", _vp), ("x_lo", _vp), ("g_hi", _vp), ("g_lo", _vp), ("dw", _vp),
                ("splits", C.c_int32), ("precision", C.c_int32)]


class Tail(C.Structure):
    _fields_ = [("size", C.c_uint32),
                ("BT", C.c_int32), ("T", C.c_int32), ("C", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("H", C.c_int32), ("W", C.c_int32),
                ("teacher_logits", _vp), ("y", _vp), ("affine", _vp), ("affine_inv", _vp), ("running_conf", _vp),
                ("training", C.c_int32), ("discount", C.c_int32),
                ("beta", C.c_float), ("stat_momentum", C.c_float), ("conf_upper", C.c_float), ("conf_lower", C.c_float),
                ("probs", _vp), ("pooled", _vp), ("part_sums", _vp), ("peaks", _vp),
                ("conf", _vp), ("idx", _vp), ("labels", _vp), ("conf_mean", _vp), ("thresholds", _vp), ("refined", _vp), ("phase", C.c_int32), ("pool_mode", C.c_int32)]


class Loss(C.Structure):
    _fields_ = [("size", C.c_uint32),
                ("BT", C.c_int32), ("C", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("logits", _vp), ("y", _vp), ("labels", _vp), ("conf_mean", _vp), ("running_conf", _vp),
                ("focal_p", C.c_float),
                ("losses", _vp), ("scratch", _vp),
                ("grad_scale", C.c_float),
                ("dlogits", _vp), ("grad_px", _vp), ("grad_rows", _vp)]


class PrepItem(C.Structure):
    _fields_ = [("w", _vp), ("gamma", _vp), ("beta", _vp), ("mean", _vp), ("var", _vp), ("conv_bias", _vp),
                ("scale", _vp), ("shift", _vp), ("wf_hi", _vp), ("wf_lo", _vp), ("wt_hi", _vp), ("wt_lo", _vp),
                ("K", C.c_int32), ("C", C.c_int32), ("R", C.c_int32), ("S", C.c_int32), ("Kf", C.c_int32), ("Kt", C.c_int32),
                ("fold_wf", C.c_int32), ("reserved", C.c_int32)]


class FinalizeItem(C.Structure):
    _fields_ = [("dwraw", _vp), ("w", _vp), ("scale", _vp), ("mean", _vp), ("var", _vp), ("dbeta", _vp),
                ("dw", _vp), ("dgamma", _vp), ("conv_bias", _vp), ("dbias", _vp), ("dbeta_out", _vp),
                ("K", C.c_int32), ("C", C.c_int32), ("RS", C.c_int32), ("splits", C.c_int32)]


def dptr(t):
    """raw device address (int) of a tensor or None -- for descriptor tables that live on the device"""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def item_table(items, blocks, device):
    """(items_dev, block_begin_dev, n_items, total_blocks) for the batched multi-layer kernels"""
    arr = (type(items[0]) * len(items))(*items)
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    begin, tot = [], 0
    for b in blocks:
        begin.append(tot); tot += int(b)
    begin.append(tot)
    return raw, torch.tensor(begin, dtype=torch.int32, device=device), len(items), tot


# optional per-launch CUDA-event profiling of the GEMM kernels (used by bench.py's roofline leg only)
_prof = None
# True while that profiling step runs: the two-stream modes (teacher || student forward, wgrad || dgrad) are switched off so that
# every kernel is timed alone on the launching stream -- events around a launch that overlaps another stream's kernel would
# charge it with the other kernel's time
SERIALIZE = False


def profile_begin():
    global _prof, SERIALIZE
    _prof = []
    SERIALIZE = True


def profile_end():
    global _prof, SERIALIZE
    rec, _prof = _prof, None
    SERIALIZE = False
    torch.cuda.synchronize()
    return [(kind, flops, e0.elapsed_time(e1)) for (kind, flops, e0, e1) in rec]


def _prof_wrap(kind, flops, fn):
    if _prof is None:
        return fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    _prof.append((kind, flops, e0, e1))
    return r


# Precision policy: "parity" = bf16x3 everywhere (default; what every golden test and the headline bench
# run), "fast_bwd" = single-pass bf16 for the data / filter gradient GEMMs only, "fast" = single-pass bf16 everywhere.  The
# engines announce the phase they are in; the wrappers below translate (policy, phase) into the descriptor's precision field.
PRECISION = os.environ.get("SACB_PRECISION", "parity")
assert PRECISION in ("parity", "fast_bwd", "fast"), "SACB_PRECISION must be parity, fast_bwd or fast"
_phase = "fwd"


def set_phase(phase):
    global _phase
    _phase = phase


def _precision():
    return 1 if (PRECISION == "fast" or (PRECISION == "fast_bwd" and _phase == "bwd")) else 0


def conv_out_hw(H, W, R, stride, dil, pad):
    return ((H + 2 * pad - (R - 1) * dil - 1) // stride + 1, (W + 2 * pad - (R - 1) * dil - 1) // stride + 1)


def conv_gemm(x_hi, x_lo, wt_hi, wt_lo, geom, *, k_valid=None, scale=None, shift=None, add_f32=None, add_hi=None,
              add_lo=None, mask_hi=None, relu=False, out_hi=None, out_lo=None, out_f32=None, out_nchw=None, colsum=None,
              unit_scale=False):
    """geom = (N, H, W, C, K, R, stride, dil, pad).  ``unit_scale``: the caller promises scale == 1 (BN scale folded into the weights)"""
    N, H, W, Cc, K, R, s, d, p = geom
    P, Q = conv_out_hw(H, W, R, s, d, p)
    desc = ConvGemm(C.sizeof(ConvGemm), N, H, W, Cc, K, K if k_valid is None else k_valid, R, R, s, d, p, P, Q,
                    ptr(x_hi), ptr(x_lo), ptr(wt_hi), ptr(wt_lo), ptr(scale), ptr(shift), ptr(add_f32), ptr(add_hi),
                    ptr(add_lo), ptr(mask_hi), 1 if relu else 0, ptr(out_hi), ptr(out_lo), ptr(out_f32), ptr(out_nchw), ptr(colsum),
                    _precision(), 1 if unit_scale else 0)
    # mirrors the dispatch in csrc/sacb_gemm.cu (sacb_conv_gemm)
    pair = os.environ.get("SACB_PAIR", "1") != "0" and K % 256 == 0
    kind = "conv_gemm_pair<256x256>" if pair else "conv_gemm<%d>" % (128 if K % 128 == 0 else (64 if K % 64 == 0 else 32))
    flops = 2.0 * N * P * Q * (K if k_valid is None else k_valid) * Cc * R * R
    _prof_wrap(kind, flops, lambda: check(lib().sacb_conv_gemm(C.byref(desc), stream()), "sacb_conv_gemm"))


def conv_wgrad(x_hi, x_lo, g_hi, g_lo, dw, geom, *, k_valid=None, splits=0):
    """dw: fp32 tensor, or a callable n_elems -> tensor used to size the split-K partial planes.
    Returns (dw, n_splits): dw holds n_splits planes [k_valid][R*R][C] that sacb_wgrad_finalize sums."""
    N, H, W, Cc, K, R, s, d, p = geom
    P, Q = conv_out_hw(H, W, R, s, d, p)
    kv = K if k_valid is None else k_valid
    desc = ConvWgrad(C.sizeof(ConvWgrad), N, H, W, Cc, K, kv, R, R, s, d, p, P, Q,
                     ptr(x_hi), ptr(x_lo), ptr(g_hi), ptr(g_lo), None, splits, _precision())
    n = lib().sacb_conv_wgrad_splits(C.byref(desc))
    if n <= 0:
        check(n if n < 0 else -1, "sacb_conv_wgrad_splits")
    need = n * kv * R * R * Cc
    if callable(dw):
        dw = dw(need)
    assert dw.numel() >= need, "split-K workspace too small: %d < %d" % (dw.numel(), need)
    desc.dw = ptr(dw)
    flops = 2.0 * N * P * Q * kv * Cc * R * R
    pair = os.environ.get("SACB_PAIR", "1") != "0" and K % 256 == 0 and Cc % 256 == 0 and not (kv <= 64 and Cc >= 128)
    _prof_wrap("conv_wgrad_pair<256x256>" if pair else "conv_wgrad", flops, lambda: check(lib().sacb_conv_wgrad(C.byref(desc), stream()), "sacb_conv_wgrad"))
    return dw, n
