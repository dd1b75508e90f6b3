"""Peer-memory plumbing for the fused gradient all-reduce + SGD kernel (``sacb_allreduce_sgd``, include/sacb.h).

``P2PContext(backbone, world, rank)`` (collective over the default process group):
  * re-homes the backbone's flat parameter buffer and flat gradient buffer into ``sacb_symm_alloc`` memory
    (plain cudaMalloc, so it can be exported; ``nn.Parameter.data`` stay views, nothing else changes),
  * exchanges CUDA IPC handles of (params, grads, flags) with ``torch.distributed.all_gather_object`` and maps every
    peer's buffers into this process (NVLink peer access is enabled lazily by the driver),
  * ``allreduce_sgd(...)`` launches the ONE kernel that replaces DDP's gradient all-reduce (train.py:104,232) and
    ``optim.step()`` (train.py:233).  No NCCL call is involved, so the whole training step -- including the exchange --
    can be captured in a CUDA graph at any world size.

torch is used for process-group plumbing and as the owner of ordinary tensors only; the shared buffers are owned by
libsac_b200.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib as L

MAX_WORLD = 8


class AllreduceSgd(C.Structure):
    _fields_ = [("size", C.c_uint32), ("world", C.c_int32), ("rank", C.c_int32),
                ("grads", C.c_void_p), ("params", C.c_void_p), ("flags", C.c_void_p),
                ("mom", C.c_void_p), ("seg_ranges", C.c_void_p), ("seg_lr", C.c_void_p), ("seg_wd", C.c_void_p),
                ("nseg", C.c_int32), ("n", C.c_int64), ("momentum", C.c_float), ("first_step", C.c_int32),
                ("mc_grads", C.c_void_p), ("mc_params", C.c_void_p)]


class SymmBuffer(object):
    """device memory from sacb_symm_alloc, viewable as a torch tensor (zero-copy) and exportable as a CUDA IPC handle"""

    def __init__(self, nbytes, device):
        self.nbytes, self.device = int(nbytes), torch.device(device)
        p = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().sacb_symm_alloc(C.c_size_t(self.nbytes), C.byref(p)), "sacb_symm_alloc")
        self.ptr = p.value

    def tensor(self, dtype, numel):
        item = torch.empty((), dtype=dtype).element_size()
        assert numel * item <= self.nbytes
        typestr = {torch.float32: "<f4", torch.uint32: "<u4", torch.int32: "<i4"}[dtype]
        holder = type("_SymmView", (), {})()
        holder.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": typestr, "data": (self.ptr, False), "version": 2}
        holder._owner = self                        # keeps the allocation alive as long as the tensor lives
        t = torch.as_tensor(holder, device=self.device)
        assert t.data_ptr() == self.ptr, "torch copied the symmetric buffer instead of aliasing it"
        return t

    def handle(self):
        h = (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            L.check(L.lib().sacb_ipc_export(C.c_void_p(self.ptr), h), "sacb_ipc_export")
        return bytes(h)


def _import(handle, device):
    p = C.c_void_p()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    with torch.cuda.device(device):
        L.check(L.lib().sacb_ipc_import(buf, C.byref(p)), "sacb_ipc_import")
    return p.value


class TorchSymmBuffer(object):
    """NVLS variant: the buffer comes from torch.distributed._symmetric_memory (CUDA VMM allocation bound to a multicast
    object by ``rendezvous``), which provides the peers' unicast pointers AND the multicast pointer; torch is the plumbing
    here exactly as it is for process groups.  Same duck type as SymmBuffer where P2PContext needs it."""

    def __init__(self, numel, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.device = torch.device(device)
        try:                                          # older torch versions want the group enabled explicitly
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        self.t = symm_mem.empty(int(numel), dtype=torch.float32, device=self.device)
        self.t.zero_()
        self.hdl = symm_mem.rendezvous(self.t, group)
        self.ptr = self.t.data_ptr()
        if not int(self.hdl.multicast_ptr):
            raise L.SacbError("torch symmetric memory reports no multicast (NVLS) support on this system")

    def tensor(self, dtype, numel):
        assert dtype == torch.float32 and numel <= self.t.numel()
        return self.t[:numel]

    def peer_ptrs(self):
        return [int(p) for p in self.hdl.buffer_ptrs]

    def multicast_ptr(self):
        return int(self.hdl.multicast_ptr)


class P2PContext(object):
    def __init__(self, backbone, world, rank, device, nvls=False):
        assert 1 <= world <= MAX_WORLD, "sacb_allreduce_sgd supports up to %d ranks (one NVSwitch domain)" % MAX_WORLD
        self.world, self.rank, self.device = world, rank, torch.device(device)
        self.backbone = backbone
        self.nvls = bool(nvls) and world > 1
        lib = L.lib()
        flat = backbone.ensure_flat(self.device)
        n = flat.total
        assert n % 4 == 0
        self.n = n
        if self.nvls:
            # SACB_NVLS=1 (verified on 2 and 8 B200s in round 2, profiles/r2*_nvls*): multimem.ld_reduce / multimem.st through the
            # NVSwitch instead of W peer loads / W peer stores per element
            grp = dist.group.WORLD
            self._bufs = dict(params=TorchSymmBuffer(n, device, grp), grads=TorchSymmBuffer(n, device, grp),
                              flags=SymmBuffer(4 * lib.sacb_p2p_flag_words(), device))
        else:
            self._bufs = dict(params=SymmBuffer(4 * n, device), grads=SymmBuffer(4 * n, device),
                              flags=SymmBuffer(4 * lib.sacb_p2p_flag_words(), device))
        # ---- re-home the flat buffers (parameter objects keep their identity; values are carried over)
        p_sym = self._bufs["params"].tensor(torch.float32, n)
        p_sym.copy_(flat.buf)
        flat.buf = p_sym
        tensors = backbone._named_tensors()
        for k, _, _ in flat.entries:
            tensors[k].data = flat.view(k)
        self.grad_sym = self._bufs["grads"].tensor(torch.float32, n)
        backbone._grad.buf = self.grad_sym
        for p in backbone._params:
            p.grad = None
        backbone.mark_dirty()
        torch.cuda.synchronize(self.device)
        # ---- exchange IPC handles, map the peers
        mine = {k: b.handle() for k, b in self._bufs.items() if isinstance(b, SymmBuffer)}
        if world > 1:
            allh = [None] * world
            dist.all_gather_object(allh, mine)
        else:
            allh = [mine]
        self._ptrs = {k: [] for k in self._bufs}
        err = None
        try:
            for r in range(world):
                for k in mine:
                    self._ptrs[k].append(self._bufs[k].ptr if r == rank else _import(allh[r][k], self.device))
            for k, b in self._bufs.items():     # NVLS: torch's rendezvous already mapped the peers
                if isinstance(b, TorchSymmBuffer):
                    self._ptrs[k] = b.peer_ptrs()
                    assert len(self._ptrs[k]) == world and self._ptrs[k][rank] == b.ptr
        except L.SacbError as e:                # e.g. no peer access between two GPUs
            err = e
        if world > 1:
            # agreement + barrier in one: every rank has mapped every peer before the first kernel touches them, and
            # either all ranks switch to the fused kernel or none does
            ok = torch.tensor([0.0 if err is not None else 1.0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok) == 0.0:
                raise L.SacbError("peer-memory mapping failed on at least one rank: %r" % (err,))
        elif err is not None:
            raise err
        arr = C.c_void_p * world
        self._arrays = {k: arr(*v) for k, v in self._ptrs.items()}

    def allreduce_sgd(self, grad_buf, mom, ranges, lr, wd, nseg, momentum, first_step):
        """all ranks: params -= lr * sgd_momentum(mean over ranks of grad) ; ``grad_buf`` is this rank's flat gradient"""
        if grad_buf.data_ptr() != self.grad_sym.data_ptr():
            self.grad_sym.copy_(grad_buf)       # a backward pass that landed in the alternate buffer (joint step)
        d = AllreduceSgd(C.sizeof(AllreduceSgd), self.world, self.rank,
                         C.cast(self._arrays["grads"], C.c_void_p), C.cast(self._arrays["params"], C.c_void_p),
                         C.cast(self._arrays["flags"], C.c_void_p), L.ptr(mom), L.ptr(ranges), L.ptr(lr), L.ptr(wd),
                         nseg, self.n, momentum, 1 if first_step else 0,
                         C.c_void_p(self._bufs["grads"].multicast_ptr()) if self.nvls else None,
                         C.c_void_p(self._bufs["params"].multicast_ptr()) if self.nvls else None)
        L.check(L.lib().sacb_allreduce_sgd(C.byref(d), L.stream()), "sacb_allreduce_sgd")
