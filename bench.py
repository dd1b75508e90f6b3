#!/usr/bin/env python
"""Benchmark of the SAC target training step (BASELINE.json metric: target-crops/sec, 512x512, K=3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one ``Trainer._step_target(train=True)`` with TARGET_ONLY semantics
(/root/reference/train.py:211-250) on ``configs[1]`` of BASELINE.json: ResNet-101 DeepLabv2,
8 view-groups x K=3 crops of 512x512 per GPU (weak scaling: every rank gets its own 8 groups).
Prints ONE JSON line (rank 0).  ``--impl reference`` times the oracle port of the reference's own
CPU path on the host cores (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TFLOP_PER_CROP = 1.506          # 4 x 376.52 GFLOP (teacher fwd + student fwd + dgrad + wgrad)
NUM_GROUPS, GROUP_SIZE, CROP = 8, 3, (512, 512)
METRIC = "target-crops/sec (512x512, K=3)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.power = []
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n for n in dir(nv) if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason")}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if isinstance(bit, int) and bit and (r & bit) and "None" not in n and "All" not in n:
                        self.reasons.add(n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
                time.sleep(0.1)
        except Exception as e:      # clocks are evidence, not a dependency of the measurement
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def summary(self):
        s = sorted(self.samples)
        p = sorted(self.power)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w_median": p[len(p) // 2] if p else None, "power_w_max": p[-1] if p else None, "samples": len(s)}


ARCH_YAML = {"resnet101": "configs/deeplabv2_resnet101_train.yaml", "vgg16": "configs/deeplabv2_vgg16_train.yaml",
             "fcn": "configs/fcn_vgg16_train.yaml"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_name(arch, groups, group_size, crop):
    return "%s SAC target step (teacher fwd + tail + student fwd/bwd + grad all-reduce + SGD), %d groups x K=%d crops %dx%d per GPU" % (
        {"resnet101": "ResNet-101 DeepLabv2", "vgg16": "VGG-16 DeepLabv2", "fcn": "VGG-16 FCN-8s"}[arch], groups, group_size, crop[0], crop[1])


class ReferenceStep(object):
    """One ``Trainer._step_target(train=True)`` (/root/reference/train.py:211-233, TRAIN.TARGET_ONLY) of the reference on the
    host CPU.  kind = "reference": the UNMODIFIED reference modules from baseline/_ref (installed by baseline/install_ref.py:
    ``models.get_model`` -> ``SAC.forward``, ``net.parameter_groups`` + ``torch.optim.SGD`` as base_trainer.get_optim builds it);
    kind = "port": oracle/sac_oracle.py, only when baseline/_ref is absent."""

    def __init__(self, arch, group_size, crop, threads):
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):       # the reference prints progress chatter; stdout carries the JSON line only
            self._init(arch, group_size, crop, threads)

    def step(self, batch):
        """returns self_ce"""
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):
            return self._step(batch)

    def _init(self, arch, group_size, crop, threads):
        import torch
        from da_sac_b200 import synth
        self.torch, self.K = torch, group_size
        torch.set_num_threads(threads)
        sd = {"resnet101": lambda: synth.make_backbone_params(seed=123), "vgg16": lambda: synth.make_vgg16_params(seed=321),
              "fcn": lambda: synth.make_fcn_params(seed=213)}[arch]()
        ref = os.path.join(ROOT, "baseline", "_ref")
        self.i = 0
        if os.path.isfile(os.path.join(ref, "MANIFEST.json")):
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            import install_ref
            install_ref.verify()                       # the files are the reference's, byte for byte
            sys.path.insert(0, ref)
            from core.config import cfg, cfg_from_file, cfg_from_list
            cfg_from_file(os.path.join(ref, ARCH_YAML[arch]))
            cfg_from_list(["TRAIN.GROUP_SIZE", str(group_size), "DATASET.CROP_SIZE", "(%d,%d)" % tuple(crop), "MODEL.INIT_MODEL", "",
                           "TRAIN.TARGET_ONLY", "True"])
            from models import get_model
            self.kind, self.cfg = "reference", cfg
            self.net = get_model(cfg.MODEL, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
            self.net.backbone.load_state_dict(sd, strict=True)
            self.net.train()
            self.optim = torch.optim.SGD(self.net.parameter_groups(cfg.MODEL.LR, cfg.MODEL.WEIGHT_DECAY), lr=cfg.MODEL.LR,
                                         momentum=cfg.MODEL.MOMENTUM, weight_decay=cfg.MODEL.WEIGHT_DECAY)
            self.what = "unmodified reference modules (baseline/_ref: models.get_model -> SAC.forward, torch.optim.SGD), torch %s CPU fp32" % torch.__version__
        else:
            assert arch == "resnet101", "baseline/_ref is not installed and the oracle port's step covers ResNet-101 only"
            from oracle import sac_oracle as O
            self.kind, self.O = "port", O
            self.mcfg = synth.ModelCfg()
            self.student = O.as_leaf_params(sd)
            self.teacher = {k: v.detach().clone() for k, v in self.student.items()}
            self.optim = torch.optim.SGD(O.parameter_groups(self.student, self.mcfg.LR, self.mcfg.WEIGHT_DECAY), momentum=self.mcfg.MOMENTUM)
            self.rc = torch.full((19,), self.mcfg.THRESHOLD_BETA)
            self.what = "oracle/sac_oracle.py (port; baseline/_ref not installed), torch %s CPU fp32" % torch.__version__

    def _step(self, batch):
        if self.kind == "port":
            losses, _, self.rc = self.O.sac_target_step(self.student, self.teacher, self.rc, batch, self.K, self.mcfg, optim=self.optim)
            return float(losses["self_ce"])
        cfg = self.cfg
        x, y, x2, A, Ai = [t.clone() for t in batch]
        update_teacher = self.i % cfg.MODEL.NET_MOMENTUM_ITER == 0                  # train.py:294
        losses, _ = self.net(x, y, x2, A, Ai, use_teacher=True, update_teacher=update_teacher, T=cfg.TRAIN.GROUP_SIZE)
        self.optim.zero_grad()                                                       # train.py:227-228
        (cfg.MODEL.LR_TARGET * losses["self_ce"].mean()).backward()                  # train.py:231-232
        self.optim.step()                                                            # train.py:233
        self.i += 1
        return float(losses["self_ce"].mean().item())                               # train.py:243-246


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU path on all host cores, on this arm's config / metric / unit.  Every step is a
    bounded sample of the workload -- ONE view-group (K crops) instead of the per-GPU batch of ``args.groups`` groups; crops/s
    on the CPU does not depend on the batch -- so that ``--steps K --warmup W`` finishes within minutes."""
    if rank != 0:
        return
    from da_sac_b200 import synth
    cores = host_cores()
    ref = ReferenceStep(args.arch, GROUP_SIZE, CROP, cores)
    sample_groups = 1
    batch = synth.make_target_batch(sample_groups, GROUP_SIZE, CROP, seed=0)
    times, ce = [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ce = ref.step(batch)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    crops = sample_groups * GROUP_SIZE * len(times)
    v = crops / total
    sample = "%d timed steps (%d warm-up) of %d view-group x K=%d crops %dx%d incl. SGD: %s, %d threads" % (
        len(times), args.warmup, sample_groups, GROUP_SIZE, CROP[0], CROP[1], ref.what, cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": SCALING,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.arch, args.groups, GROUP_SIZE, CROP),
                       "baseline_config": args.config if args.config is not None else 1,
                       "global_batch_crops": args.groups * GROUP_SIZE * world, "parallelism": "dp%d" % world,
                       "reference_sample": "each step = %d of the %d view-groups per GPU (bounded CPU sample)" % (sample_groups, args.groups)},
            "self_ce_last": ce,
            "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": v, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_reference_cuda(args, rank):
    """Second baseline of SURVEY.md 8(d): the same reference modules (oracle port = functional restatement of the reference's
    PyTorch code) executed by stock PyTorch on THIS GPU -- ATen / cuDNN kernels, cudnn.benchmark = True as train.py:40 sets it,
    TF32 convolutions as PyTorch's default allows -- on the full configs[1] batch.  None of this repo's kernels run here."""
    if rank != 0:
        return
    import torch
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    cfg = synth.ModelCfg()
    sd = synth.make_backbone_params(seed=123)
    student = O.as_leaf_params({k: v.to(dev) for k, v in sd.items()})
    teacher = {k: v.detach().clone() for k, v in student.items()}
    optim = torch.optim.SGD(O.parameter_groups(student, cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    rc = torch.full((19,), cfg.THRESHOLD_BETA, device=dev)
    batch = tuple(t.to(dev) for t in synth.make_target_batch(args.groups, GROUP_SIZE, CROP, seed=0))
    crops = args.groups * GROUP_SIZE

    def timed(tf32, warm, steps):
        nonlocal rc
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        for _ in range(warm):
            _, _, rc = O.sac_target_step(student, teacher, rc, batch, GROUP_SIZE, cfg, optim=optim)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            _, _, rc = O.sac_target_step(student, teacher, rc, batch, GROUP_SIZE, cfg, optim=optim)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"crops_per_s": crops / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "warmup": warm}

    r = timed(True, args.warmup, args.steps)
    line = {"impl": "reference", "device": "cuda", "metric": METRIC, "value": r["crops_per_s"], "unit": "crops/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 convolutions (PyTorch default: cudnn.allow_tf32 = True), fp32 elsewhere",
            "data": "synthetic",
            "config": {"workload": "ResNet-101 DeepLabv2 SAC target step, %d groups x K=%d crops %dx%d, stock PyTorch %s + cuDNN %s on the GPU (oracle port of the reference modules, cudnn.benchmark = True)"
                                   % (args.groups, GROUP_SIZE, CROP[0], CROP[1], torch.__version__, torch.backends.cudnn.version())},
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if args.ref_strict_fp32:
        line["strict_fp32"] = timed(False, 1, max(1, args.steps // 2))
    print(json.dumps(line), flush=True)


# BASELINE.json configs[i] -> (arch, view-groups in the whole job, K, crop, GPUs the config is stated on, scaling).
# configs[1] is the bench line (weak scaling: 8 groups per GPU whatever N); the others fix the GLOBAL batch, so
# `--config i --gpus N` gives every rank groups/N of it (strong scaling over N).  configs[0] is the CPU plumbing case.
PRESETS = {
    0: ("vgg16", 1, 1, (256, 256), 1, "strong"),
    1: ("resnet101", None, 3, (512, 512), 1, "weak"),
    2: ("resnet101", 32, 3, (512, 512), 8, "strong"),
    3: ("fcn", 16, 4, (640, 640), 4, "strong"),
    4: ("resnet101", 8, 6, (1024, 1024), 8, "strong"),
}
SCALING = "weak"


def main():
    global GROUP_SIZE, CROP, TFLOP_PER_CROP, METRIC, SCALING
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=None, choices=sorted(PRESETS),
                    help="BASELINE.json configs[i]: 1 = the bench line (default); 2/3/4 fix the global batch (32x512^2 K=3 / FCN 16x640^2 "
                         "K=4 / 8x1024^2 K=6) and split it over --gpus")
    ap.add_argument("--groups", type=int, default=None, help="view-groups per GPU (default: configs[1]: 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--arch", default=None, choices=["resnet101", "vgg16", "fcn"])
    ap.add_argument("--group-size", type=int, default=None)
    ap.add_argument("--crop", type=int, nargs=2, default=None)
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-p2p", action="store_true", help="N>1: NCCL all-reduce + SGD instead of the fused peer-memory kernel")
    ap.add_argument("--no-exchange-check", action="store_true", help="N>1: skip the bit-exactness check of the fused exchange")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference: cpu = the reference arm of the contract; cuda = stock PyTorch/cuDNN on the GPU (second baseline)")
    ap.add_argument("--ref-strict-fp32", action="store_true", help="--ref-device cuda: also time with cudnn.allow_tf32 = False")
    ap.add_argument("--also-fast", action="store_true",
                    help="after the parity-mode measurement, re-capture and time the step in the fast precision modes "
                         "(fast_bwd, fast) and report them under 'fast_modes' (separate numbers, never the headline)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    preset = PRESETS[args.config if args.config is not None else 1]
    if args.arch is None: args.arch = preset[0]
    if args.group_size is None: args.group_size = preset[2]
    if args.crop is None: args.crop = list(preset[3])
    if args.groups is None:
        if preset[1] is None:
            args.groups = NUM_GROUPS
        else:
            assert preset[1] % world == 0, "configs[%d] has %d view-groups: they do not split over %d GPUs" % (args.config, preset[1], world)
            args.groups = preset[1] // world
            SCALING = preset[5]
    default_cfg = (args.arch, args.group_size, tuple(args.crop)) == ("resnet101", 3, (512, 512))
    if not default_cfg:
        GROUP_SIZE, CROP = args.group_size, tuple(args.crop)
        METRIC = "target-crops/sec (%dx%d, K=%d)" % (CROP[0], CROP[1], GROUP_SIZE)
        # forward GFLOP per crop from SURVEY.md 8(a)/appendix D (probe of the reference modules), scaled by area otherwise
        known = {("resnet101", 512): 376.52, ("resnet101", 1024): 1483.27, ("vgg16", 512): 325.55, ("vgg16", 256): 81.39, ("fcn", 640): 346.34}
        base = {"resnet101": (376.52, 512), "vgg16": (325.55, 512), "fcn": (346.34, 640)}[args.arch]
        gf = known.get((args.arch, CROP[0])) if CROP[0] == CROP[1] else None
        TFLOP_PER_CROP = 4e-3 * (gf if gf is not None else base[0] * CROP[0] * CROP[1] / float(base[1] ** 2))

    if args.impl == "reference" and args.ref_device == "cuda":
        return run_reference_cuda(args, rank)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from da_sac_b200 import lib as L, synth
    from da_sac_b200.models import get_model
    from da_sac_b200.trainer import TargetStepper, allreduce_mean_

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.lib()        # fail loudly if the CUDA extension is missing

    cfg = {"resnet101": synth.ModelCfg, "vgg16": synth.ModelCfgVGG16, "fcn": synth.ModelCfgFCN}[args.arch]()
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the model constructors print what the reference's print; stdout carries the JSON line only
        net = get_model(cfg, rank, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    net.backbone.load_state_dict({"resnet101": lambda: synth.make_backbone_params(seed=123), "vgg16": lambda: synth.make_vgg16_params(seed=321),
                                  "fcn": lambda: synth.make_fcn_params(seed=213)}[args.arch]())
    net.to(dev).train()
    stepper = TargetStepper(net, cfg, GROUP_SIZE, dev)
    exchange = "none (1 GPU)"
    if world > 1:
        exchange = "NCCL all-reduce of the flat gradient + SGD kernel"
        if not args.no_p2p:
            # gradient mean + SGD + weight broadcast as ONE kernel over NVLink peer memory (no NCCL on the data path)
            ok = torch.ones(1, device=dev)
            try:
                stepper.enable_p2p()
            except Exception as e:          # peer access unavailable: every rank must agree before changing path
                print("[bench] rank %d: peer-memory setup failed (%r); using NCCL" % (rank, e), file=sys.stderr, flush=True)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok) > 0:
                exchange = "fused peer-memory all-reduce + SGD kernel (sacb_allreduce_sgd%s, no NCCL)" % (", NVLS multimem" if stepper.optim.p2p.nvls else "")
            else:
                stepper.optim.p2p = None
    host = stepper.stage_host(synth.make_target_batch(args.groups, GROUP_SIZE, CROP, seed=rank))
    dev_batch = stepper.h2d(host)
    crops_per_step = args.groups * GROUP_SIZE * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}

    def run(n, e2e):
        for _ in range(n):
            if e2e:
                # pinned host -> device copy of EVERY step's inputs (issued on a copy stream while the previous step
                # computes, as a pin_memory DataLoader would) + loss scalars -> host
                last["losses"] = stepper.step(host, read_losses=True, prefetch_next=host)
            elif stepper._graph is not None:
                last["losses"] = stepper.step(dev_batch, read_losses=False)    # inputs are copied into the graph's static buffers
            else:
                b = tuple(t.clone() if i == 1 else t for i, t in enumerate(dev_batch))   # y is mutated in place
                last["losses"] = stepper.step(b, read_losses=False)

    def timed(n, e2e):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        run(n, e2e)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)              # max over ranks
        return float(ms)

    def note(msg):
        if rank == 0:
            print("[bench] " + msg, file=sys.stderr, flush=True)

    note("first eager step (teacher init, workspace allocation)")
    run(1, False)

    # ---- N > 1: what does the fused exchange compute?  One more optimiser step from the state every rank holds now:
    # (a) the fused peer-memory kernel (reduce-scatter + SGD + all-gather), (b) from the SAME parameters / momentum / local
    # gradients: NCCL all-reduce (sum, / world: DDP, train.py:104) followed by the single-GPU sacb_sgd kernel.  At world 2 the
    # mean (a+b)/2 has one rounding whatever the order, so the two must agree bit for bit; at world > 2 NCCL's ring / tree
    # order differs from the kernel's fixed 0..W-1 order, so the bound is a few ulp of the update.  All replicas must end up
    # with identical bits either way (each element is reduced by exactly one rank).
    exchange_check = None
    if world > 1 and stepper.optim.p2p is not None and not args.no_exchange_check:
        bb, opt = net.backbone, stepper.optim
        if opt._built is None: opt._build()
        b_ = opt._built
        p0, m0, g0, steps0 = bb._flat.buf.clone(), b_["mom"].clone(), bb._grad.buf.clone(), opt.steps
        barrier()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        opt.step()                                               # (a) fused
        ev1.record()
        torch.cuda.synchronize()
        barrier()
        ex_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        dist.all_reduce(ex_ms, op=dist.ReduceOp.MAX)
        p_fused, m_fused = bb._flat.buf.clone(), b_["mom"].clone()
        # (b) NCCL + sacb_sgd from the same state.  The fused kernel keeps the momentum SHARDED (ZeRO-1: a rank only ever touches
        # the slice it owns, the rest of its buffer stays zero), so the full momentum is the sum of the ranks' buffers.
        m_full = m0.clone()
        dist.all_reduce(m_full)
        bb._flat.buf.copy_(p0); b_["mom"].copy_(m_full); opt.steps = steps0
        gref = g0.clone()
        allreduce_mean_(gref)
        import ctypes as C
        L.check(L.lib().sacb_sgd(L.ptr(bb._flat.buf), L.ptr(gref), L.ptr(b_["mom"]), L.ptr(b_["ranges"]), L.ptr(b_["lr"]), L.ptr(b_["wd"]),
                                 b_["n"], C.c_float(opt.momentum), 1 if steps0 == 0 else 0, L.stream()), "sacb_sgd")
        torch.cuda.synchronize()
        p_nccl = bb._flat.buf.clone()
        upd = (p_nccl - p0).double()
        diff = (p_fused.double() - p_nccl.double())
        stats = torch.tensor([float((p_fused != p_nccl).sum()), float(diff.norm() / upd.norm().clamp_min(1e-300))], device=dev, dtype=torch.float64)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        # replicas: checksum of the raw bit patterns must be the same on every rank
        bits = p_fused.view(torch.int32).to(torch.int64)
        cks = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
        lo, hi = cks.clone(), cks.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_equal = bool((lo == hi).all())
        bit_exact = stats[0].item() == 0
        n_opt = sum(r1 - r0 for r0, r1 in zip(b_["ranges_host"][0::2], b_["ranges_host"][1::2]))     # optimiser elements
        link_bytes = (world - 1) / world * n_opt * 4 * (1.0 / world if opt.p2p.nvls else 1.0)
        exchange_check = {"vs_nccl_allreduce_plus_sgd": "bit-exact" if bit_exact else "rel-L2 of the update %.3e (%d elements differ)" % (stats[1].item(), int(stats[0].item())),
                          "replicas_equal": replicas_equal, "update_rel_l2": stats[1].item(), "elements": int(bits.numel()),
                          # one launch of the fused kernel right after a barrier (ranks aligned): reduce-scatter + SGD + all-gather of
                          # the whole flat buffer; bytes a rank pulls from / pushes to its peers over NVLink, each way
                          "fused_kernel_ms": float(ex_ms), "nvlink_bytes_per_rank_each_way": link_bytes,
                          "nvlink_gbs_each_way": link_bytes / (float(ex_ms) * 1e-3) / 1e9}
        note("exchange check: %s" % exchange_check)
        assert replicas_equal, "replicas diverged after the fused exchange"
        # W > 2: measured 4e-7 ... 6e-7 of the update (profiles/r2h_*, r2i_*: fp32 summation order of 4 / 8 addends); a wrong exchange is O(1)
        assert bit_exact or (world > 2 and stats[1].item() < 1e-5), "fused exchange differs from NCCL all-reduce + SGD: %s" % exchange_check
        # continue from the fused result (identical on all ranks) with the fused kernel's own (sharded) momentum
        bb._flat.buf.copy_(p_fused); b_["mom"].copy_(m_fused); opt.steps = steps0 + 1
        bb.mark_dirty()
        barrier()

    # the fused exchange kernel makes the whole step NCCL-free, so it is captured at any N; with NCCL: eager launches
    use_graph = (not args.no_graph) and (world == 1 or stepper.optim.p2p is not None)
    if use_graph:
        note("capturing the steady-state step (and the teacher-update step) into CUDA graphs")
        stepper.capture(dev_batch)
    note("warm-up x%d" % args.warmup)
    run(args.warmup, False)
    def measure():
        # "update_teacher EMA every 100th step amortised".  The timed region is positioned so that exactly one
        # teacher-update step (iter % NET_MOMENTUM_ITER == 0) falls inside it: 1 in `steps` instead of 1 in 100 -- conservative.
        stepper.iter = cfg.NET_MOMENTUM_ITER - args.steps // 2
        smp = ClockSampler(local_rank)
        smp.start()
        l0 = stepper.launches
        t = timed(args.steps, False)
        n = stepper.launches - l0
        smp.stop_flag = True
        smp.join(timeout=2)
        return t, n, smp

    def throttled(smp):
        """timing rules: hardware / thermal slowdown, or SM clocks far below max with no stated reason (a leftover clock lock),
        invalidate a run; a software power cap is normal for this workload and only noted"""
        c = smp.summary()
        bad = any(r in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown") for r in c["reasons"])
        stuck = bool(c["sm_mhz"] and c["sm_max_mhz"] and c["sm_mhz"] < 0.6 * c["sm_max_mhz"] and not c["reasons"])
        flag = torch.tensor([1.0 if (bad or stuck) else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)          # every rank must take the same decision
        return float(flag) > 0

    note("timed region x%d" % args.steps)
    ms, launches, sampler = measure()
    remeasured = False
    if throttled(sampler):
        note("clock throttling seen during the timed region (%s): measuring once more" % sampler.summary()["reasons"])
        time.sleep(5.0)
        ms, launches, sampler = measure()
        remeasured = True
    self_ce_dev = float(last["losses"]["self_ce"])               # loss of the last timed step (device-resident run)
    run(1, True)
    stepper.iter = cfg.NET_MOMENTUM_ITER - args.steps // 2
    ms_e2e = timed(args.steps, True)
    self_ce_e2e = float(last["losses"]["self_ce"])

    replicas_after = None
    if world > 1:
        bits = net.backbone._flat.buf.view(torch.int32).to(torch.int64)
        ck = bits.sum().reshape(1)
        lo, hi = ck.clone(), ck.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_after = bool((lo == hi).all())                  # after warm-up + 2 x steps optimiser steps

    fast_modes = None
    if args.also_fast and L.PRECISION == "parity":
        # plain bf16 is the "fast mode, reported separately".  Same step, same batch; only the precision field of
        # the conv descriptors changes, so the graph is re-captured per mode.  The parity numbers above are already taken.
        fast_modes = {}
        for mode in ("fast_bwd", "fast"):
            L.PRECISION = mode
            stepper.drop_graphs()
            run(1, False)
            if use_graph:
                stepper.capture(dev_batch)
            run(args.warmup, False)
            stepper.iter = cfg.NET_MOMENTUM_ITER - args.steps // 2
            ms_f = timed(args.steps, False)
            fast_modes[mode] = {"value": crops_per_step * args.steps / (ms_f / 1e3), "unit": "crops/s", "ms_per_step": ms_f / args.steps,
                                "self_ce_last": float(last["losses"]["self_ce"])}
        L.PRECISION = "parity"
        stepper.drop_graphs()

    value = crops_per_step * args.steps / (ms / 1e3)
    e2e = crops_per_step * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host)
    pk = peaks()

    # ---- roofline of the dominant kernel, measured live with CUDA events on the launching stream (one extra step)
    graphs = stepper.suspend_graphs()                   # per-launch events need eager launches
    stepper.iter = 1
    L.profile_begin()
    run(1, False)
    prof = L.profile_end()
    stepper.resume_graphs(graphs)
    by = {}
    for kind, flops, t in prof:
        a = by.setdefault(kind, [0.0, 0.0, 0]); a[0] += flops; a[1] += t; a[2] += 1
    dom = max(by, key=lambda k: by[k][1])
    achieved = by[dom][0] / (by[dom][1] * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel: taken from the committed ncu --set full
    # capture that profiles/roofline_traffic.json names (written by profiles/ncu_summary.py from the .ncu-rep of this round);
    # null when there is no capture for this configuration's dominant kernel
    traffic, traffic_note = None, None
    tj = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if default_cfg and os.path.isfile(tj):
        t_ = json.load(open(tj))
        if dom.startswith(t_.get("kernel_prefix", "?")):
            traffic, traffic_note = t_["dram_bytes_per_launch"], t_["note"]
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["sustained"], "traffic": traffic, "traffic_note": traffic_note,
                "launches": by[dom][2],
                "share_of_step": by[dom][1] / (ms / args.steps),
                "peak_source": "%s bf16 sustained (MEASURED_PEAKS.json); precision mode '%s': in the parity mode the kernel issues 3 bf16 MMAs per algorithmic MAC (bf16x3 split), so frac <= 1/3 by construction" % (pk["src"], L.PRECISION),
                "step_tensor_frac": value / world * TFLOP_PER_CROP / pk["sustained"],
                "kernels": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms": v[1], "launches": v[2]} for k, v in by.items()}}

    line = {"metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
            "dtype": {"parity": "bf16x3 (bf16 hi/lo split operands, fp32 TMEM accumulation; fp32-equivalent)",
                      "fast_bwd": "bf16x3 forward (fp32-equivalent logits / pseudo labels), single-pass bf16 gradient GEMMs (SACB_PRECISION=fast_bwd)",
                      "fast": "bf16 single-pass everywhere, fp32 TMEM accumulation (SACB_PRECISION=fast; NOT the parity mode: logits ~1e-2)"}[L.PRECISION],
            "data": "synthetic",
            "config": {"workload": workload_name(args.arch, args.groups, GROUP_SIZE, CROP),
                       "baseline_config": args.config if args.config is not None else 1,
                       "global_batch_crops": crops_per_step, "parallelism": "dp%d" % world, "gradient_exchange": exchange,
                       "precision_mode": L.PRECISION,
                       "teacher_update_steps_in_timed_region": 1,
                       "l2": "inputs larger than L2 (>20 GB of activations per step)",
                       "launch": "CUDA graph replay of the whole step" if use_graph else "eager launches"},
            "e2e": {"value": e2e, "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": ms_e2e / args.steps},
            "self_ce_last": {"device_resident": self_ce_dev, "e2e": self_ce_e2e},
            "gpu_launches": launches, "clocks": dict(sampler.summary(), remeasured=remeasured), "roofline": roofline,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if exchange_check is not None:
        line["exchange_check"] = exchange_check
        line["replicas_equal_after_timed_steps"] = replicas_after
    if fast_modes is not None:
        line["fast_modes"] = fast_modes

    if rank == 0 and world == 1 and not args.no_cpu_baseline and default_cfg:
        # bounded sample (about 15-20 s of CPU work): 1 warm-up + 3 timed steps of ONE view-group (K crops) each
        cores = host_cores()
        ref = ReferenceStep(args.arch, GROUP_SIZE, CROP, cores)
        cb = synth.make_target_batch(1, GROUP_SIZE, CROP, seed=0)
        dts = []
        for i in range(4):
            t0 = time.perf_counter()
            ref.step(cb)
            if i > 0: dts.append(time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": GROUP_SIZE * len(dts) / sum(dts), "unit": "crops/s", "cores": cores, "kind": ref.kind,
                                "sample": "%d timed steps (1 warm-up) of 1 view-group x K=%d crops %dx%d incl. SGD: %s, %d threads" % (len(dts), GROUP_SIZE, CROP[0], CROP[1], ref.what, cores)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
