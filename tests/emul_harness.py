"""TEST INFRASTRUCTURE: runs the unmodified GPU parity tests in the GPU-less container.

`emulated_gpu()` (a context manager) makes the product's Python layer talk to tests/cpu_emul/_build/libsacb_emul.so instead of
libsac_b200.so: the streaming kernels execute from their real CUDA source under the host emulation (cuda_emul.h), the two
tensor-core entry points through the checker's plain-loop model (gemm_model.cpp).  "cuda" tensors are created on the CPU
(a TorchFunctionMode rewrites device arguments), so a GPU test function can be called as it is.

What a green run here proves: the host-side schedules (engine.py, engine_abn.py, models/) and the streaming kernels'
indexing / arithmetic.  What it does not prove: anything about the tcgen05 GEMM kernels, streams, graphs, or speed -- the GPU
suite stays the parity gate.  The product never loads this library (da_sac_b200/lib.py refuses it)."""
import contextlib
import ctypes as C
import os
import shutil
import subprocess

import torch
from torch.overrides import TorchFunctionMode

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "cpu_emul")
SO = os.environ.get("SACB_EMUL_SO") or os.path.join(EMUL, "_build", "libsacb_emul.so")     # an ASan build: see cpu_emul/Makefile

_lib = None
_libs = {}


def available():
    return shutil.which("g++") is not None and shutil.which("make") is not None


def emul_lib():
    global _lib
    if _lib is None:
        if not os.environ.get("SACB_EMUL_SO"):
            r = subprocess.run(["make", "-C", EMUL], capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
        lib = C.CDLL(SO)
        assert lib.sacb_emul_marker() == 1
        lib.sacb_last_error.restype = C.c_char_p
        lib.sacb_launch_count.restype = C.c_int64
        for f in ("sacb_tail_part_sums_elems", "sacb_tail_probs_elems", "sacb_tail_pooled_elems"):
            getattr(lib, f).restype = C.c_size_t
        lib.sacb_bn_moments_partial_elems.restype = C.c_size_t
        lib.sacb_bn_moments_partial_elems.argtypes = [C.c_int64, C.c_int]
        _lib = lib
    return _lib


def emul_lib_full():
    """every kernel from its real source: the streaming units AND sacb_gemm.cu (on the primitive model of cuda_emul_tc.h)"""
    if "full" not in _libs:
        emul_lib()                                            # builds all libraries
        lib = C.CDLL(os.path.join(os.path.dirname(SO), "libsacb_emul_full.so"))
        assert lib.sacb_emul_marker() == 1
        lib.sacb_last_error.restype = C.c_char_p
        lib.sacb_launch_count.restype = C.c_int64
        lib.sacb_emul_last_kernel.restype = C.c_char_p
        for f in ("sacb_tail_part_sums_elems", "sacb_tail_probs_elems", "sacb_tail_pooled_elems"):
            getattr(lib, f).restype = C.c_size_t
        lib.sacb_bn_moments_partial_elems.restype = C.c_size_t
        lib.sacb_bn_moments_partial_elems.argtypes = [C.c_int64, C.c_int]
        _libs["full"] = lib
    return _libs["full"]


def _is_cuda(d):
    try:
        return d is not None and not isinstance(d, (bool, int)) and torch.device(d).type == "cuda"
    except (TypeError, RuntimeError):
        return False


class _CudaToCpu(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if func in (torch.Tensor.cuda,):
            return args[0]
        if _is_cuda(kwargs.get("device")):
            kwargs["device"] = "cpu"
        if func is torch.Tensor.to and len(args) > 1 and isinstance(args[1], (str, torch.device)) and _is_cuda(args[1]):
            args = (args[0], "cpu") + tuple(args[2:])
        if func is torch.Tensor.pin_memory:
            return args[0]
        return func(*args, **kwargs)


class _FakeStream(object):
    """torch.cuda.Stream stand-in: the emulation runs everything in program order, so fork / join are no-ops.  What this lets
    a test execute is the HOST logic of a multi-stream path (which engine, which buffers, which order), not its concurrency."""
    cuda_stream = 0

    def wait_stream(self, other):
        pass

    def __init__(self, *a, **k):
        pass

    def wait_event(self, ev):
        pass

    def record_event(self, ev=None):
        return ev

    def synchronize(self):
        pass


class _FakeEvent(object):
    """torch.cuda.Event stand-in: records host time, so elapsed_time is host milliseconds (meaningless as a measurement)"""

    def __init__(self, enable_timing=False, **kw):
        self.t = None

    def record(self, stream=None):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)

    def synchronize(self):
        pass

    def query(self):
        return True


@contextlib.contextmanager
def emulated_gpu(full=False):
    """full=False: tensor-core entry points through the formula model (fast); full=True: through the real kernel source"""
    from da_sac_b200 import lib as L
    lib = emul_lib_full() if full else emul_lib()

    def ptr(t):
        if t is None:
            return None
        assert t.device.type == "cpu" and t.is_contiguous()
        return C.c_void_p(t.data_ptr())

    from da_sac_b200 import p2p as P2P
    saved = {(L, k): getattr(L, k) for k in ("lib", "stream", "ptr", "dptr", "on_device")}
    saved[(P2P.SymmBuffer, "tensor")] = P2P.SymmBuffer.tensor
    saved[(torch.cuda, "device")] = torch.cuda.device

    def symm_tensor(self, dtype, numel):
        """host view of a sacb_symm_alloc buffer (the product builds it through __cuda_array_interface__)"""
        import numpy as np
        npdt = {torch.float32: np.float32, torch.uint32: np.uint32, torch.int32: np.int32}[dtype]
        assert numel * np.dtype(npdt).itemsize <= self.nbytes
        a = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(self.nbytes,))[:numel * np.dtype(npdt).itemsize].view(npdt)
        t = torch.from_numpy(a)
        if dtype == torch.uint32:
            t = t.view(torch.uint32) if t.dtype != torch.uint32 else t
        assert t.data_ptr() == self.ptr
        t._symm_owner = self
        return t
    for k in ("Stream", "current_stream", "stream", "Event", "set_device", "max_memory_allocated"):
        saved[(torch.cuda, k)] = getattr(torch.cuda, k)
    saved[(torch.cuda, "synchronize")] = torch.cuda.synchronize
    saved[(torch.cuda, "is_current_stream_capturing")] = torch.cuda.is_current_stream_capturing
    L.lib = lambda: lib
    L.stream = lambda: C.c_void_p(0)
    L.ptr = ptr
    L.dptr = lambda t: None if t is None else t.data_ptr()
    L.on_device = lambda t: True
    P2P.SymmBuffer.tensor = symm_tensor
    torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    torch.cuda.Stream = _FakeStream
    torch.cuda.current_stream = lambda *a, **k: _FakeStream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.Event = _FakeEvent
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.max_memory_allocated = lambda *a, **k: 0
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.is_current_stream_capturing = lambda: False
    try:
        with _CudaToCpu():
            yield lib
    finally:
        for (obj, k), v in saved.items():
            setattr(obj, k, v)
