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


def run_reference(args, rank):
    """the reference's own CPU implementation of the path: oracle port (torch CPU fp32, all host threads)"""
    if rank != 0:
        return
    import torch
    from da_sac_b200 import synth
    from oracle import sac_oracle as O
    cores = min(os.cpu_count() or 1, 32)         # oneDNN convs on 65x65 maps stop scaling past ~32 threads
    torch.set_num_threads(cores)
    cfg = synth.ModelCfg()
    sd = synth.make_backbone_params(seed=123)
    student = O.as_leaf_params(sd)
    teacher = {k: v.detach().clone() for k, v in student.items()}
    optim = torch.optim.SGD(O.parameter_groups(student, cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
    rc = torch.full((19,), cfg.THRESHOLD_BETA)
    groups = 1                                   # bounded sample: 1 group x K=3 crops of 512x512 per step
    batch = synth.make_target_batch(groups, GROUP_SIZE, CROP, seed=0)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, _, rc = O.sac_target_step(student, teacher, rc, batch, GROUP_SIZE, cfg, optim=optim)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    crops = groups * GROUP_SIZE * len(times)
    v = crops / total
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ResNet-101 DeepLabv2 SAC target step, %d group x K=%d crops %dx%d per step (bounded CPU sample of configs[1])" % (groups, GROUP_SIZE, CROP[0], CROP[1])},
            "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cores, "kind": "port",
                             "sample": "%d steps of 1 group x K=3 crops 512x512 (oracle/sac_oracle.py, torch CPU fp32)" % len(times)},
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


def main():
    global GROUP_SIZE, CROP, TFLOP_PER_CROP, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--groups", type=int, default=NUM_GROUPS, help="view-groups per GPU (default: configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # the other BASELINE.json configs (parity-test cases, not the bench line): e.g. configs[3] per-GPU shard
    # `--arch fcn --groups 4 --group-size 4 --crop 640 640`, configs[4] `--groups 1 --group-size 6 --crop 1024 1024`
    ap.add_argument("--arch", default="resnet101", choices=["resnet101", "vgg16", "fcn"])
    ap.add_argument("--group-size", type=int, default=GROUP_SIZE)
    ap.add_argument("--crop", type=int, nargs=2, default=list(CROP))
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-p2p", action="store_true", help="N>1: NCCL all-reduce + SGD instead of the fused peer-memory kernel")
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
    if args.impl == "reference" and args.ref_device == "cuda":
        return run_reference_cuda(args, rank)
    if args.impl == "reference":
        if args.steps > 3: args.steps = 3
        if args.warmup > 1: args.warmup = 1
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from da_sac_b200 import lib as L, synth
    from da_sac_b200.models import get_model
    from da_sac_b200.trainer import TargetStepper

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    default_cfg = (args.arch, args.group_size, tuple(args.crop)) == ("resnet101", 3, (512, 512))
    if not default_cfg:
        GROUP_SIZE, CROP = args.group_size, tuple(args.crop)
        METRIC = "target-crops/sec (%dx%d, K=%d)" % (CROP[0], CROP[1], GROUP_SIZE)
        # forward GFLOP per crop from SURVEY.md 8(a)/appendix D (probe of the reference modules), scaled by area otherwise
        known = {("resnet101", 512): 376.52, ("resnet101", 1024): 1483.27, ("vgg16", 512): 325.55, ("vgg16", 256): 81.39, ("fcn", 640): 346.34}
        base = {"resnet101": (376.52, 512), "vgg16": (325.55, 512), "fcn": (346.34, 640)}[args.arch]
        gf = known.get((args.arch, CROP[0])) if CROP[0] == CROP[1] else None
        TFLOP_PER_CROP = 4e-3 * (gf if gf is not None else base[0] * CROP[0] * CROP[1] / float(base[1] ** 2))
        args.no_cpu_baseline = True
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.lib()        # fail loudly if the CUDA extension is missing

    cfg = {"resnet101": synth.ModelCfg, "vgg16": synth.ModelCfgVGG16, "fcn": synth.ModelCfgFCN}[args.arch]()
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
                exchange = "fused peer-memory all-reduce + SGD kernel (sacb_allreduce_sgd, no NCCL)"
            else:
                stepper.optim.p2p = None
    host = stepper.stage_host(synth.make_target_batch(args.groups, GROUP_SIZE, CROP, seed=rank))
    dev_batch = stepper.h2d(host)
    crops_per_step = args.groups * GROUP_SIZE * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n, e2e):
        for _ in range(n):
            if e2e:
                # pinned host -> device copy of EVERY step's inputs (issued on a copy stream while the previous step
                # computes, as a pin_memory DataLoader would) + loss scalars -> host
                stepper.step(host, read_losses=True, prefetch_next=host)
            elif stepper._graph is not None:
                stepper.step(dev_batch, read_losses=False)         # inputs are copied into the graph's static buffers
            else:
                b = tuple(t.clone() if i == 1 else t for i, t in enumerate(dev_batch))   # y is mutated in place
                stepper.step(b, read_losses=False)

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
    # the fused exchange kernel makes the whole step NCCL-free, so it is captured at any N; with NCCL: eager launches
    use_graph = (not args.no_graph) and (world == 1 or stepper.optim.p2p is not None)
    if use_graph:
        note("capturing the steady-state step into a CUDA graph")
        stepper.capture(dev_batch)
    note("warm-up x%d" % args.warmup)
    run(args.warmup, False)
    def measure():
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
    run(1, True)
    ms_e2e = timed(args.steps, True)

    fast_modes = None
    if args.also_fast and L.PRECISION == "parity":
        # plain bf16 is the "fast mode, reported separately".  Same step, same batch; only the precision field of
        # the conv descriptors changes, so the graph is re-captured per mode.  The parity numbers above are already taken.
        fast_modes = {}
        for mode in ("fast_bwd", "fast"):
            L.PRECISION = mode
            stepper._graph = None
            run(1, False)
            if use_graph:
                stepper.capture(dev_batch)
            run(args.warmup, False)
            ms_f = timed(args.steps, False)
            fast_modes[mode] = {"value": crops_per_step * args.steps / (ms_f / 1e3), "unit": "crops/s", "ms_per_step": ms_f / args.steps}
        L.PRECISION = "parity"
        stepper._graph = None

    value = crops_per_step * args.steps / (ms / 1e3)
    e2e = crops_per_step * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host)
    pk = peaks()

    # ---- roofline of the dominant kernel, measured live with CUDA events on the launching stream (one extra step)
    graph, stepper._graph = stepper._graph, None        # per-launch events need eager launches
    L.profile_begin()
    run(1, False)
    prof = L.profile_end()
    stepper._graph = graph
    by = {}
    for kind, flops, t in prof:
        a = by.setdefault(kind, [0.0, 0.0, 0]); a[0] += flops; a[1] += t; a[2] += 1
    dom = max(by, key=lambda k: by[k][1])
    achieved = by[dom][0] / (by[dom][1] * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel on its most frequent shape
    # (3x3 d2 256->256, 23 layers x 3 passes), from the committed ncu --set full capture profiles/ncu_gemm_pair_r1k.txt
    traffic = 106.904832e6 + 62.850816e6 if (dom.startswith("conv_gemm_pair") and default_cfg) else None
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["sustained"], "traffic": traffic,
                "traffic_note": "bytes per launch on the 3x3 d2 256->256 layer (algorithmic activations in+out 208 MB, weights 2.4 MB); per-launch FLOPs there: 119.6 GFLOP",
                "launches": by[dom][2],
                "share_of_step": by[dom][1] / (ms / args.steps),
                "peak_source": "%s bf16 sustained (MEASURED_PEAKS.json); precision mode '%s': in the parity mode the kernel issues 3 bf16 MMAs per algorithmic MAC (bf16x3 split), so frac <= 1/3 by construction" % (pk["src"], L.PRECISION),
                "step_tensor_frac": value / world * TFLOP_PER_CROP / pk["sustained"],
                "kernels": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms": v[1], "launches": v[2]} for k, v in by.items()}}

    line = {"metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"parity": "bf16x3 (bf16 hi/lo split operands, fp32 TMEM accumulation; fp32-equivalent)",
                      "fast_bwd": "bf16x3 forward (fp32-equivalent logits / pseudo labels), single-pass bf16 gradient GEMMs (SACB_PRECISION=fast_bwd)",
                      "fast": "bf16 single-pass everywhere, fp32 TMEM accumulation (SACB_PRECISION=fast; NOT the parity mode: logits ~1e-2)"}[L.PRECISION],
            "data": "synthetic",
            "config": {"workload": "%s SAC target step (teacher fwd + tail + student fwd/bwd + grad all-reduce + SGD), %d groups x K=%d crops %dx%d per GPU" % ({"resnet101": "ResNet-101 DeepLabv2", "vgg16": "VGG-16 DeepLabv2", "fcn": "VGG-16 FCN-8s"}[args.arch], args.groups, GROUP_SIZE, CROP[0], CROP[1]),
                       "global_batch_crops": crops_per_step, "parallelism": "dp%d" % world, "gradient_exchange": exchange,
                       "precision_mode": L.PRECISION,
                       "l2": "inputs larger than L2 (>20 GB of activations per step)",
                       "launch": "CUDA graph replay of the whole step" if use_graph else "eager launches"},
            "e2e": {"value": e2e, "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": dict(sampler.summary(), remeasured=remeasured), "roofline": roofline}
    if fast_modes is not None:
        line["fast_modes"] = fast_modes

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import sac_oracle as O
        cores = min(os.cpu_count() or 1, 32)
        torch.set_num_threads(cores)
        sd = synth.make_backbone_params(seed=123)
        student = O.as_leaf_params(sd)
        teacher = {k: v.detach().clone() for k, v in student.items()}
        rc = torch.full((19,), cfg.THRESHOLD_BETA)
        cb = synth.make_target_batch(1, GROUP_SIZE, CROP, seed=0)
        optim = torch.optim.SGD(O.parameter_groups(student, cfg.LR, cfg.WEIGHT_DECAY), momentum=cfg.MOMENTUM)
        # bounded sample (about 10-20 s of CPU work): 1 warm-up + 3 timed steps of ONE view-group (3 crops) each
        dts = []
        for i in range(4):
            t0 = time.perf_counter()
            _, _, rc = O.sac_target_step(student, teacher, rc, cb, GROUP_SIZE, cfg, optim=optim)
            if i > 0: dts.append(time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": GROUP_SIZE * len(dts) / sum(dts), "unit": "crops/s", "cores": cores, "kind": "port",
                                "sample": "%d timed steps (1 warm-up) of 1 group x K=3 crops 512x512 incl. SGD (oracle/sac_oracle.py, torch CPU fp32, %d threads)" % (len(dts), cores)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
