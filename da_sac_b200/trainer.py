"""Target-step driver: the B200-side mirror of ``Trainer._step_target`` / ``_prep_batch``
(/root/reference/train.py:157-250) with TRAIN.TARGET_ONLY semantics.

``TargetStepper.step(batch_target, update_teacher)``:
    H2D (optional, pinned) -> SAC.forward (teacher fwd, tail, student fwd, fused loss)
    -> zero_grad -> (LR_TARGET * self_ce).backward() -> gradient all-reduce (mean over ranks, NCCL)
    -> SGD (one multi-tensor kernel over the flat parameter buffer) -> loss scalars.

Data parallelism follows the reference: whole view-groups are partitioned across ranks
(``shard_groups``), every rank holds a full replica, the only exchange is the gradient all-reduce."""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib as L


def shard_groups(num_groups, world_size, rank):
    """indices of the view-groups rank ``rank`` owns (datasets/__init__.py:64: NUM_GROUPS // ngpus per GPU)"""
    if num_groups % world_size != 0:
        raise ValueError("NUM_GROUPS=%d does not split over %d ranks" % (num_groups, world_size))
    per = num_groups // world_size
    return list(range(rank * per, (rank + 1) * per))


def sync_bn_sums(sums, count, group=None):
    """Cross-rank part of nn.SyncBatchNorm in the ABN baseline (deeplabv2.py:15,183; train.py:104): the per-channel sums a
    rank computed over its local batch ([k, C] float64 -- forward: sum z, sum z^2; backward: sum g, sum g*xhat) and its
    element count are summed over the ranks, so that every rank normalises with the statistics of the GLOBAL batch.
    torch's SyncBatchNorm all-gathers (mean, invstd, count) per rank and recombines them; summing raw moments is the
    same statistic in one all-reduce.  In place; returns (sums, total_count).  World size 1: no communication."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        buf = torch.cat([sums.reshape(-1), count.reshape(-1).to(sums.dtype)])
        dist.all_reduce(buf, group=group)
        sums.copy_(buf[:-1].view_as(sums))
        count = buf[-1:].clone()
    return sums, float(count.reshape(-1)[0])


def shard_batch(batch, group_size, world_size, rank):
    """slice a flattened [G*T, ...] target batch to this rank's groups (train.py:186-187 early-out path)"""
    G = batch[0].shape[0] // group_size
    idx = shard_groups(G, world_size, rank)
    lo, hi = idx[0] * group_size, (idx[-1] + 1) * group_size
    return tuple(t[lo:hi] for t in batch)


def fractional_subgroup(rank, local_batch, group_size):
    """(first rank, number of ranks) of the set of ranks that hold the views of this rank's view-group when a group
    does not fit one GPU (sac.py:204-214: ``stride = T // B``; ``index = stride * (rank * B // T)``)."""
    stride = max(1, group_size // local_batch)
    return stride * (rank * local_batch // group_size), stride


def prep_batch(tensor, num_groups, group_size, world_size, rank, device=None):
    """``Trainer._prep_batch`` (train.py:157-209): ``tensor`` is this rank's loader output [B,T,...].  If at least one
    whole group fits a GPU the batch is just flattened to [B*T,...]; otherwise every rank's batch is all-gathered and
    this rank keeps its ``N*L/world`` consecutive views of the group it shares with its neighbours."""
    N, L_ = num_groups, group_size
    if (N * L_) % world_size != 0:
        raise ValueError("Batch size does not fit world size")                    # train.py:179
    per_gpu = N * L_ // world_size
    if device is not None:
        tensor = tensor.to(device, non_blocking=True)
    if per_gpu >= L_:
        return tensor.flatten(0, 1)                                               # train.py:186-187
    T = tensor.size(1)
    assert T == L_, "Loaded sequence is incorrect {} vs. {}".format(T, L_)       # train.py:193
    out_list = [torch.empty_like(tensor) for _ in range(world_size)]
    dist.all_gather(out_list, tensor.contiguous())
    batch_index = rank * per_gpu
    index0, index1 = batch_index // T, batch_index % T
    return out_list[index0].flatten(0, 1)[index1:index1 + per_gpu]


def allreduce_mean_(buf):
    """DDP semantics for gradients: sum over ranks, divide by world size (train.py:104). In place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf)
        buf.div_(dist.get_world_size())
    return buf


class FusedSGD(object):
    """torch.optim.SGD(momentum, weight_decay per group) semantics (base_trainer.py:61-66) as ONE kernel over
    the backbone's flat parameter / gradient buffers. ``param_groups`` come from ``net.parameter_groups``."""

    def __init__(self, backbone, param_groups, momentum=0.9):
        self.backbone, self.momentum = backbone, momentum
        self.param_groups = param_groups
        self._built = None
        self.steps = 0
        self.p2p = None        # p2p.P2PContext: step() then also performs the gradient all-reduce (one fused kernel)
        self.frozen = False    # set by TargetStepper.capture(): device tensors are referenced by a CUDA graph from then on

    def _build(self, mom=None):
        flat = self.backbone._flat
        ranges, lr, wd = self._tables()
        dev = flat.buf.device
        self._built = dict(flat=flat, flat_ptr=flat.buf.data_ptr(), ranges_host=ranges,
                           ranges=torch.tensor(ranges, dtype=torch.int64, device=dev),
                           lr=torch.tensor(lr, dtype=torch.float32, device=dev),
                           wd=torch.tensor(wd, dtype=torch.float32, device=dev), n=len(lr),
                           mom=torch.zeros_like(flat.buf) if mom is None else mom)

    def _tables(self):
        """(ranges, lr, wd) host lists over the optimiser tensors in flat-buffer order"""
        bb = self.backbone
        flat = bb._flat
        by_id = {}
        for g in self.param_groups:
            for p in g["params"]:
                by_id[id(p)] = (g["lr"], g["weight_decay"])
        tensors = dict(bb.named_parameters())
        ranges, lr, wd = [], [], []
        for key, _, is_p in flat.entries:
            if not is_p: continue
            p = tensors[key]
            if id(p) not in by_id: continue
            o, n, _, _ = flat.offs[key]
            ranges += [o, o + n]; lr.append(by_id[id(p)][0]); wd.append(by_id[id(p)][1])
        return ranges, lr, wd

    def set_lr(self, param_groups):
        """New learning rates / weight decays (the reference rebuilds nothing either: ``base_trainer`` pokes
        ``param_group["lr"]``).  The momentum buffer, the segment table and the device tensors that a captured CUDA graph
        reads through raw pointers all stay where they are; only their contents change."""
        self.param_groups = param_groups
        b = self._built
        if b is None:
            return
        ranges, lr, wd = self._tables()
        if len(lr) != b["n"] or ranges != b["ranges_host"]:
            raise ValueError("FusedSGD.set_lr: the set of optimised tensors changed; build a new FusedSGD instead")
        b["lr"].copy_(torch.tensor(lr, dtype=torch.float32), non_blocking=False)
        b["wd"].copy_(torch.tensor(wd, dtype=torch.float32), non_blocking=False)

    def zero_grad(self):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    def step(self):
        bb = self.backbone
        b = self._built
        if b is None or b["flat"] is not bb._flat or b["flat_ptr"] != bb._flat.buf.data_ptr():
            # first step, or the flat parameter buffer moved (e.g. re-homed into peer memory by enable_p2p): the layout is
            # the same, so the momentum accumulated so far is carried over
            assert not self.frozen, "the flat parameter buffer changed after the step was captured into a CUDA graph"
            mom = b["mom"] if (b is not None and b["mom"].numel() == bb._flat.buf.numel()
                               and b["mom"].device == bb._flat.buf.device) else None
            self._build(mom)
        b = self._built
        if self.p2p is not None:
            # gradient mean over ranks + SGD + weight broadcast in ONE kernel over NVLink peer memory (sacb_allreduce_sgd)
            self.p2p.allreduce_sgd(bb._grad.buf, b["mom"], b["ranges"], b["lr"], b["wd"], b["n"], self.momentum,
                                   self.steps == 0)
            self.steps += 1
            bb.mark_dirty()
            return
        L.check(L.lib().sacb_sgd(L.ptr(b["flat"].buf), L.ptr(bb._grad.buf), L.ptr(b["mom"]), L.ptr(b["ranges"]),
                                 L.ptr(b["lr"]), L.ptr(b["wd"]), b["n"], C.c_float(self.momentum),
                                 1 if self.steps == 0 else 0, L.stream()), "sacb_sgd")
        self.steps += 1
        bb.mark_dirty()


class TargetStepper(object):
    def __init__(self, net, cfg_model, group_size, device, lr=None, weight_decay=None):
        self.net, self.cfg, self.T, self.device = net, cfg_model, group_size, device
        net.backbone.ensure_flat(device); net.slow_net.ensure_flat(device)
        groups = net.parameter_groups(cfg_model.LR if lr is None else lr,
                                      cfg_model.WEIGHT_DECAY if weight_decay is None else weight_decay)
        self.optim = FusedSGD(net.backbone, groups, cfg_model.MOMENTUM)
        self.iter = 0
        self._pinned = None
        self._graph = None
        self._graph_upd = None
        self._static = None
        self._graph_losses = None
        self._graph_launches = 0
        self.launches = 0          # kernels of libsac_b200 launched (eagerly or through graph replays) by step()

    def enable_p2p(self, world=None, rank=None):
        """Switch the gradient exchange from NCCL all-reduce + SGD to the fused peer-memory kernel (collective call:
        every rank must make it, before ``capture``).  With it the whole step has no NCCL call inside, so the CUDA-graph
        replay path works at any world size."""
        from .p2p import P2PContext
        if world is None:
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            rank = dist.get_rank() if world > 1 else 0
        assert self._graph is None, "enable_p2p() must precede capture()"
        import os
        self.optim.p2p = P2PContext(self.net.backbone, world, rank, self.device, nvls=os.environ.get("SACB_NVLS", "0") == "1")
        return self.optim.p2p

    def stage_host(self, batch):
        """pinned host copies of a batch (what a DataLoader with pin_memory hands to train.py:183)"""
        self._pinned = tuple(t.contiguous().pin_memory() for t in batch)
        return self._pinned

    def h2d(self, host_batch):
        return tuple(t.to(self.device, non_blocking=True) for t in host_batch)

    def prefetch(self, host_batch):
        """Start the host->device copy of the NEXT batch on a side stream so that it overlaps the step that is running
        (the B200-side equivalent of a DataLoader with pin_memory + ``.cuda(non_blocking=True)``; the reference copies
        synchronously at train.py:183).  ``step(host_batch)`` then picks the staged copy up."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_batch)
            self._staging_free = None
        assert all(d.shape == s_.shape and d.dtype == s_.dtype for d, s_ in zip(self._staging, host_batch)), \
            "prefetch(): batch shapes / dtypes differ from the first prefetched batch"
        cs = self._copy_stream
        if self._staging_free is not None:
            cs.wait_event(self._staging_free)               # the previous step has finished reading the staging buffers
        with torch.cuda.stream(cs):
            for d, s_ in zip(self._staging, host_batch):
                d.copy_(s_, non_blocking=True)
            self._staging_ready = torch.cuda.Event()
            self._staging_ready.record(cs)
        self._prefetched = host_batch

    def _take_prefetched(self, batch):
        """device copy of ``batch`` if ``prefetch(batch)`` is pending, else None"""
        if getattr(self, "_prefetched", None) is None or any(a is not b for a, b in zip(self._prefetched, batch)):
            return None
        self._prefetched = None
        torch.cuda.current_stream().wait_event(self._staging_ready)
        return self._staging

    def _release_staging(self):
        self._staging_free = torch.cuda.Event()
        self._staging_free.record(torch.cuda.current_stream())

    def _eager(self, batch, update_teacher):
        x, y, x2, A, Ai = batch
        losses, outs = self.net(x, y, x2, A, Ai, use_teacher=True, update_teacher=update_teacher, T=self.T)
        self.optim.zero_grad()                                                      # TARGET_ONLY (train.py:227-228)
        (self.cfg.LR_TARGET * losses["self_ce"].mean()).backward()                  # train.py:231-232
        if self.optim.p2p is None:
            allreduce_mean_(self.net.backbone._grad.buf)                            # NCCL; else fused into optim.step()
        self.optim.step()                                                           # train.py:233
        return torch.cat([losses["loss_ce"].detach(), losses["self_ce"].detach(), losses["teacher_diff"].detach()])

    def capture(self, example_batch, teacher_update_graph=True):
        """Capture the step into CUDA graphs: the whole step is a fixed kernel schedule, so replaying it removes ~900
        launches' worth of host work per step.  Two graphs: the steady-state step and (``teacher_update_graph``) the step
        that also applies the teacher EMA (every NET_MOMENTUM_ITER-th, train.py:294).  Call after at least one eager step.
        The warm-up passes that stream capture needs are real steps on ``example_batch``; student / teacher parameters,
        momentum, ``running_conf`` and the step counters are snapshotted before and restored afterwards, so capturing does
        not train."""
        assert self.iter > 0, "run an eager step first (teacher initialisation, workspace allocation)"
        bb, tn = self.net.backbone, self.net.slow_net
        self._static = tuple(t.clone() for t in example_batch)
        src = tuple(t.clone() for t in example_batch)
        if self.optim._built is None:
            self.optim._build()
        snap = (bb._flat.buf.clone(), self.optim._built["mom"].clone(), self.net.running_conf.clone(), tn._flat.buf.clone(),
                self.optim.steps, self.iter)
        self.optim.steps = max(self.optim.steps, 1)      # the captured kernel is the steady-state one (momentum buffer in use)

        def one(update):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    for d, s_ in zip(self._static, src): d.copy_(s_)
                    self._eager(self._static, update)
            torch.cuda.current_stream().wait_stream(side)
            for d, s_ in zip(self._static, src): d.copy_(s_)
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(g):
                losses = self._eager(self._static, update)
            return g, losses, L.launch_count() - n0       # kernel nodes of ours inside the graph

        self._graph, self._graph_losses, self._graph_launches = one(False)
        self._graph_upd = one(True) if teacher_update_graph else None
        self.optim.frozen = True
        bb._flat.buf.copy_(snap[0]); self.optim._built["mom"].copy_(snap[1]); self.net.running_conf.copy_(snap[2])
        tn._flat.buf.copy_(snap[3])
        bb.mark_dirty(); tn.mark_dirty()
        with torch.no_grad():
            tn._planes(False)            # the steady-state graph does not re-derive the teacher's weight planes
        self.optim.steps, self.iter = snap[4], snap[5]
        if self.optim.steps == 0:
            # a graph captured before any optimiser step would have baked first_step=1 in; the momentum buffer is zero
            # then, and buf = 0.9 * 0 + d == d, so the steady-state kernel gives the same first update
            self.optim.steps = 1
        return self._graph

    def drop_graphs(self):
        self._graph = self._graph_upd = None
        self.optim.frozen = False

    def suspend_graphs(self):
        """eager launches for a while (per-launch profiling); returns the token for ``resume_graphs``"""
        tok = (self._graph, getattr(self, "_graph_upd", None))
        self._graph = self._graph_upd = None
        return tok

    def resume_graphs(self, tok):
        self._graph, self._graph_upd = tok

    def step(self, batch, update_teacher=None, read_losses=False, prefetch_next=None):
        """one Trainer._step_target(train=True); ``batch`` may live on the device or in pinned host memory.
        ``prefetch_next``: pinned host batch of the following step, copied while this step computes."""
        if update_teacher is None:
            update_teacher = (self.iter % self.cfg.NET_MOMENTUM_ITER == 0)          # train.py:294
        staged = None if batch[0].is_cuda else self._take_prefetched(batch)
        graph = None
        if self._graph is not None:
            graph = (self._graph, self._graph_losses, self._graph_launches) if not update_teacher else getattr(self, "_graph_upd", None)
        if graph is not None:
            for d, s_ in zip(self._static, batch if staged is None else staged):
                d.copy_(s_, non_blocking=True)                                      # H2D when ``batch`` is pinned host memory
            if staged is not None:
                self._release_staging()
            graph[0].replay()
            self.launches += graph[2]
            v = graph[1].clone()                                                    # the next replay overwrites the static output
        else:
            if staged is not None:
                batch = tuple(t.clone() for t in staged)                            # y is mutated in place (sac.py:338)
                self._release_staging()
            elif not batch[0].is_cuda:
                batch = self.h2d(batch)
            n0 = L.launch_count()
            v = self._eager(batch, update_teacher)
            self.launches += L.launch_count() - n0
        if prefetch_next is not None:
            self.prefetch(prefetch_next)
        self.iter += 1
        if read_losses:
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                v = v.clone(); dist.all_reduce(v); v /= dist.get_world_size()       # train.py:243-245
            host = v.cpu()                                                          # .item() host sync (train.py:246)
            return {"loss_ce": float(host[0]), "self_ce": float(host[1]), "teacher_diff": float(host[2])}
        return {"loss_ce": v[0:1], "self_ce": v[1:2], "teacher_diff": v[2:3]}


class JointStepper(TargetStepper):
    """One iteration of ``Trainer.train_epoch`` with TRAIN.TARGET_ONLY = False (train.py:266-298): supervised source pass
    (``Trainer.step(train=True)``, train.py:119-138: zero_grad, ``loss_ce.backward()``, no optimiser step) followed by the
    target pass (``_step_target``, train.py:211-233: ``LR_TARGET * self_ce`` backward WITHOUT zero_grad, then
    ``optim.step()``).  The two backward passes land in two flat buffers that are summed by one kernel; there is ONE
    gradient all-reduce for the pair (DDP in the reference all-reduces after each backward)."""

    def step_joint(self, batch_source, batch_target, update_teacher=None):
        if update_teacher is None:
            update_teacher = (self.iter % self.cfg.NET_MOMENTUM_ITER == 0)
        bb = self.net.backbone
        image, masks_gt = batch_source
        if not image.is_cuda:
            image, masks_gt = self.h2d((image, masks_gt))
        if not batch_target[0].is_cuda:
            batch_target = self.h2d(batch_target)
        n0 = L.launch_count()
        losses_src, _ = self.net(image, masks_gt)                                   # train.py:128
        self.optim.zero_grad()                                                      # train.py:132
        losses_src["loss_ce"].mean().backward()                                     # train.py:133
        g_src = bb.hold_grad()
        self.optim.zero_grad()            # drop the per-tensor views; the held flat buffer keeps the source gradients
        x, y, x2, A, Ai = batch_target
        losses, _ = self.net(x, y, x2, A, Ai, use_teacher=True, update_teacher=update_teacher, T=self.T)
        (self.cfg.LR_TARGET * losses["self_ce"].mean()).backward()                  # train.py:231-232 (accumulates)
        assert bb._grad is not g_src
        bb._grad.buf.add_(g_src.buf)
        if self.optim.p2p is None:
            allreduce_mean_(bb._grad.buf)
        self.optim.step()                                                           # train.py:233
        self.launches += L.launch_count() - n0
        self.iter += 1
        return {"loss_ce_source": losses_src["loss_ce"].detach(), "loss_ce": losses["loss_ce"].detach(),
                "self_ce": losses["self_ce"].detach(), "teacher_diff": losses["teacher_diff"].detach()}
