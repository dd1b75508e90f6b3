"""Runs the GPU parity tests -- the SAME test functions, unmodified -- in the GPU-less container on the host emulation library
(tests/emul_harness.py: streaming kernels from their real CUDA source under tests/cpu_emul/cuda_emul.h, the two tensor-core
entry points through the checker's plain-loop model).  A green run proves the host-side schedules and the streaming kernels
against the goldens of the real reference; it says nothing about the tcgen05 kernels, which only `pytest -m gpu` on a B200 can.

Most valuable for code that has not run on a B200 yet (tests/test_abn_gpu.py): its full iteration is executed here.

The default selection keeps the CPU suite within a few minutes; SACB_EMUL_FULL=1 runs everything that can be emulated
(about 25 minutes on 8 cores, dominated by the 30-step training test)."""
import importlib
import inspect
import os

import numpy as np
import pytest
import torch

import emul_harness as E

pytestmark = pytest.mark.skipif(not E.available(), reason="no host toolchain for tests/cpu_emul")
FULL = os.environ.get("SACB_EMUL_FULL") == "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (module, test function, parametrisation ids to run by default: None = all, () = only under SACB_EMUL_FULL=1)
SELECTION = [
    ("test_conv_gpu", "test_conv_fprop_plain", None),                     # these five check gemm_model.cpp itself (vs torch fp64)
    ("test_conv_gpu", "test_conv_fprop_fused_epilogue", None),
    ("test_conv_gpu", "test_conv_fprop_mask_and_add_f32", None),
    ("test_conv_gpu", "test_conv_aspp_head_nchw", None),
    ("test_conv_gpu", "test_conv_wgrad", None),
    ("test_step_gpu", "test_backbone_forward_matches_oracle", None),
    ("test_step_gpu", "test_tail_labels_bit_exact_on_golden_teacher_logits", None),
    ("test_step_gpu", "test_two_training_steps_match_reference_golden", ()),     # default coverage: smoke() on the real source + two-stream, below
    ("test_step_gpu", "test_vgg16_config1_two_steps_match_reference_golden", ()),
    ("test_step_gpu", "test_backward_matches_reference_given_golden_pseudo_labels", ()),
    ("test_step_gpu", "test_source_pass_loss_ce_backward_matches_oracle", ()),
    ("test_step_gpu", "test_fcn8s_two_steps_match_reference_golden", ()),
    ("test_step_gpu", "test_fcn8s_dropout_train_mode_runs", ()),
    ("test_step_gpu", "test_joint_source_target_step_matches_oracle", ()),
    ("test_step_gpu", "test_training_reduces_the_loss_and_teacher_follows", ()),
    ("test_variants_gpu", "test_tail_variant_labels_and_losses", None),
    ("test_fractional_gpu", "test_whole_group_tail_unchanged_by_phase_split", None),
    ("test_augment_gpu", "test_augment_matches_oracle_and_reference_golden", None),
    ("test_ragged_gpu", "test_ragged_crop_full_step_vs_cpu_oracle", (1,)),   # odd sides 97 x 131 by default; the other two under FULL
    ("test_ragged_gpu", "test_single_view_group", ()),
    ("test_ragged_gpu", "test_every_pixel_ignored", None),
    ("test_prepare_finalize_gpu", "test_prepare_batched_planes_bit_exact", None),
    ("test_prepare_finalize_gpu", "test_wgrad_finalize_batched_bit_exact_dw", None),
    ("test_abn_gpu", "test_bn_moments_survive_large_mean", None),
    ("test_abn_gpu", "test_abn_iteration_matches_reference_golden", (0, 1)),  # never run on a B200: the reason this file exists
                                                                               # (0 resnet101, 1 vgg16 by default; 2 fcn under FULL)
]


def _cases():
    out = []
    # NB: nothing here may touch SACB_RUN_UNVERIFIED -- this module is also IMPORTED (collected, then deselected) by
    # `pytest -m gpu` on the GPU box, and the gates of the not-yet-verified GPU tests must stay shut there.  Calling a test
    # function directly ignores its module's skipif mark, so the variable is not needed.
    for mod, fn, ids in SELECTION:
        f = getattr(importlib.import_module(mod), fn)
        params = [m for m in getattr(f, "pytestmark", []) if m.name == "parametrize"]
        if not params:
            variants = [((), {})]
        else:
            assert len(params) == 1, "stacked parametrisations are not handled"
            names = [n.strip() for n in params[0].args[0].split(",")]
            variants = [((), dict(zip(names, v if len(names) > 1 else (v,)))) for v in params[0].args[1]]
        for i, (_, kw) in enumerate(variants):
            default = ids is None or i in ids
            out.append(pytest.param(mod, fn, kw, id="%s::%s[%d]" % (mod, fn, i),
                                    marks=[] if (default or FULL) else [pytest.mark.skip(reason="SACB_EMUL_FULL=1 runs it")]))
    return out


@pytest.fixture          # function scope on purpose: the two libraries must never be patched in at the same time
def emul():
    torch.set_num_threads(8)
    with E.emulated_gpu() as lib:
        yield lib


@pytest.fixture
def emul_full():
    """every kernel from its real source, the tcgen05 / TMA GEMM kernels included (tests/cpu_emul/cuda_emul_tc.h)"""
    torch.set_num_threads(8)
    with E.emulated_gpu(full=True) as lib:
        yield lib


_fixture_cache = {}


def _fixture(mod, name, request, fresh=False):
    if name == "golden":
        return np.load(os.path.join(ROOT, "tests", "golden", "sac_resnet101_tiny.npz"), allow_pickle=False)
    if name in ("monkeypatch", "tmp_path"):
        return request.getfixturevalue(name)
    if fresh:                                              # e.g. a `net` no earlier test has trained
        return getattr(mod, name).__wrapped__()
    key = (mod.__name__, name)
    if key not in _fixture_cache:                          # module-scoped fixtures of the GPU test modules (e.g. `net`)
        _fixture_cache[key] = getattr(mod, name).__wrapped__()
    return _fixture_cache[key]


@pytest.mark.parametrize("mod,fn,kw", _cases())
def test_gpu_test_on_the_emulation_library(emul, request, mod, fn, kw):
    m = importlib.import_module(mod)
    f = getattr(m, fn)
    kw = dict(kw)
    for name in inspect.signature(f).parameters:
        if name not in kw:
            kw[name] = _fixture(m, name, request)
    n0 = emul.sacb_launch_count()
    f(**kw)
    assert emul.sacb_launch_count() > n0, "the test did not reach the library"


# ---------------------------------------------------------------- the same, with NO formula model: sacb_gemm.cu's real kernels
FULL_SELECTION = [   # (module, function, parametrisation index or None, runs by default)
    ("test_step_gpu", "test_backbone_forward_matches_oracle", None, False),       # the smoke() test below covers it by default
    ("test_step_gpu", "test_two_training_steps_match_reference_golden", None, False),      # 913 launches, 77 s on 8 cores
    ("test_step_gpu", "test_vgg16_config1_two_steps_match_reference_golden", None, False),
    ("test_abn_gpu", "test_abn_iteration_matches_reference_golden", 1, False),      # vgg16
    ("test_ragged_gpu", "test_ragged_crop_full_step_vs_cpu_oracle", 1, False),      # 97 x 131: ragged M tiles, odd maps through im2col TMA
]


@pytest.mark.parametrize("mod,fn,idx,default", FULL_SELECTION)
def test_gpu_test_on_the_real_kernel_source_of_everything(request, mod, fn, idx, default):
    if not (default or FULL):
        pytest.skip("SACB_EMUL_FULL=1 runs it")
    lib = request.getfixturevalue("emul_full")
    m = importlib.import_module(mod)
    f = getattr(m, fn)
    kw = {}
    if idx is not None:
        mark = [x for x in f.pytestmark if x.name == "parametrize"][0]
        names = [n.strip() for n in mark.args[0].split(",")]
        v = mark.args[1][idx]
        kw = dict(zip(names, v if len(names) > 1 else (v,)))
    for name in inspect.signature(f).parameters:
        if name not in kw:
            kw[name] = _fixture(m, name, request, fresh=True)
    n0 = lib.sacb_launch_count()
    f(**kw)
    assert lib.sacb_launch_count() > n0
    assert b"conv_" in lib.sacb_emul_last_kernel()             # a tcgen05 kernel instantiation was launched through the primitive model


SWITCH_STEP = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import emul_harness as E, test_step_gpu as G
golden = np.load(os.path.join(%r, "tests", "golden", "sac_resnet101_tiny.npz"), allow_pickle=False)
with E.emulated_gpu(full=True) as lib:
    G.test_two_training_steps_match_reference_golden(G.net.__wrapped__(), golden)
print("OK")
'''


@pytest.mark.skipif(not FULL, reason="SACB_EMUL_FULL=1 runs it")
@pytest.mark.parametrize("switch", ["SACB_EPI_STAGED", "SACB_TAIL_SPLIT"])
def test_two_training_steps_with_an_unverified_kernel_variant_switched_on(switch):
    """the whole SAC step (every kernel from real source) with the residual-staging epilogue / the tail split in the loop, vs the
    real reference's golden; SACB_EMUL_SMS=8 so that the small problem has partial last waves"""
    import subprocess, sys
    env = dict(os.environ, SACB_EMUL_SMS="8")
    env[switch] = "1"
    r = subprocess.run([sys.executable, "-c", SWITCH_STEP % (ROOT, os.path.join(ROOT, "tests"), ROOT)], env=env, capture_output=True,
                       text=True, timeout=3000)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("arch", ["vgg16", "resnet101"])
def test_two_training_steps_with_the_two_stream_schedule(emul, monkeypatch, request, arch):
    """SACB_TWO_STREAM=1 (models/sac.py: teacher forward + tail issued on a side stream with its OWN engine, joined before the
    loss) has not run on a GPU.  Streams do not exist here -- fork / join are no-ops -- but the host logic does: the second
    engine, its workspaces and plane pools, the order of the calls.  Two full training steps must still match the golden."""
    if arch == "resnet101" and not FULL:
        pytest.skip("SACB_EMUL_FULL=1 runs it")
    from da_sac_b200.models import sac as S
    m = importlib.import_module("test_step_gpu")
    monkeypatch.setattr(S, "_TWO_STREAM", True)
    made = []
    init = S.SAC.__init__

    def spy(self, *a, **k):
        init(self, *a, **k)
        made.append(self)
    monkeypatch.setattr(S.SAC, "__init__", spy)
    if arch == "vgg16":
        m.test_vgg16_config1_two_steps_match_reference_golden()
    else:
        m.test_two_training_steps_match_reference_golden(_fixture(m, "net", request, fresh=True), _fixture(m, "golden", request))
    assert made and made[-1]._engines_teacher, "the teacher did not get its own engine: the two-stream branch was not taken"


def test_driver_smoke_entry_point_on_the_real_kernel_source(emul_full, monkeypatch):
    """__graft_entry__.smoke() -- what the driver runs on the B200 at the end of every round -- unmodified, every kernel from its
    real source: a host-side edit that would break the round-end run shows up here first"""
    import __graft_entry__ as G
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    n0 = emul_full.sacb_launch_count()
    G.smoke()
    assert emul_full.sacb_launch_count() - n0 > 300


@pytest.mark.parametrize("extra", [["--arch", "vgg16"], [], ["--also-fast"]], ids=["vgg16", "resnet101", "resnet101_also_fast"])
def test_bench_main_on_the_emulation_prints_a_complete_line(emul, monkeypatch, capsys, extra):
    """bench.py is what the driver runs at the end of every round, and parts of it were edited after the last GPU call (re-measure
    on throttling, clocks.remeasured, precision mode in the line).  Its whole main() -- argument handling, TargetStepper, staging /
    prefetch, timed regions, clock sampler, roofline table, JSON line -- executes here on a tiny configuration with stand-in
    events and streams (eager launches; the CUDA-graph capture and the default-size cpu_baseline leg are not reachable here).
    The numbers mean nothing; the keys and the absence of a Python error do."""
    import json
    import runpy
    if "--arch" not in extra and not FULL:
        pytest.skip("SACB_EMUL_FULL=1 runs it")
    argv = ["bench.py", "--groups", "1", "--group-size", "2", "--crop", "64", "64", "--steps", "2", "--warmup", "3", "--no-graph"] + extra
    monkeypatch.setattr("sys.argv", argv)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert key in line, key
    assert line["steps"] == 2 and line["warmup"] == 3 and line["n_gpus"] == 1 and line["gpu_launches"] > 100
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and line["e2e"]["h2d_bytes_per_step"] > 0
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert "remeasured" in line["clocks"] and "workload" in line["config"] and line["config"]["precision_mode"] == "parity"
    if "--also-fast" in extra:
        assert set(line["fast_modes"]) == {"fast_bwd", "fast"}


def test_target_stepper_with_the_fused_peer_memory_exchange(emul):
    """TargetStepper.enable_p2p() -- the path bench.py takes at N > 1 -- at world 1 on the emulation: P2PContext re-homes the flat
    parameter / gradient buffers into sacb_symm_alloc memory, exports / imports the handles, and every step ends in
    allreduce_sgd_kernel instead of sacb_sgd.  p2p.py was edited after the last multi-GPU run (NVLS plumbing): two training steps
    must leave exactly the parameters FusedSGD leaves."""
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    from da_sac_b200.trainer import TargetStepper
    batch = synth.make_target_batch(1, 2, (64, 64), seed=0)
    res, epochs = [], None
    for use_p2p in (False, True):
        cfg = synth.ModelCfgVGG16()
        net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
        net.backbone.load_state_dict(synth.make_vgg16_params(seed=321))
        net.train()
        st = TargetStepper(net, cfg, 2, torch.device("cpu"))
        ctx = st.enable_p2p(1, 0) if use_p2p else None
        for _ in range(2):
            st.step(tuple(t.clone() for t in batch), read_losses=True)
        res.append(torch.cat([q.detach().reshape(-1).clone() for q in net.backbone.parameters()]))
        if ctx is not None:
            assert st.optim.p2p is ctx and not ctx.nvls
            epochs = int(ctx._bufs["flags"].tensor(torch.int32, emul.sacb_p2p_flag_words())[16])      # FLAG_EPOCH
    assert epochs == 2, "the fused exchange kernel did not run once per step"
    # (bit-identical on the real kernel source; the formula model's column sums are added in a run-dependent order)
    assert ((res[0] - res[1]).abs().max() / res[0].abs().max()).item() < 1e-5


def test_fused_sgd_set_lr_keeps_momentum_and_device_tensors(emul):
    """``FusedSGD.set_lr`` (learning-rate schedule between steps): the momentum buffer must survive and the device tensors a
    captured CUDA graph reads through raw pointers (lr, wd, segment table, momentum) must stay where they are -- only their
    contents change.  Checked against torch.optim.SGD taking the same two steps with the same change of learning rate."""
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    from da_sac_b200.trainer import FusedSGD
    cfg = synth.ModelCfgVGG16()
    net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    net.backbone.load_state_dict(synth.make_vgg16_params(seed=321))
    bb = net.backbone
    bb.ensure_flat(torch.device("cpu"))
    opt = FusedSGD(bb, net.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY), cfg.MOMENTUM)
    ref_params = [p.detach().clone().requires_grad_(True) for p in bb.parameters()]
    groups = net.parameter_groups(cfg.LR, cfg.WEIGHT_DECAY)
    index = {id(p): i for i, p in enumerate(bb.parameters())}
    ref_groups = [dict(params=[ref_params[index[id(p)]] for p in g["params"]], lr=g["lr"], weight_decay=g["weight_decay"]) for g in groups]
    ref = torch.optim.SGD(ref_groups, momentum=cfg.MOMENTUM)
    gen = torch.Generator().manual_seed(3)
    ptrs = None
    for step in range(3):
        g = torch.randn(bb._flat.total, generator=gen) * 1e-2
        bb._grad.buf.copy_(g)
        for k, p in zip(bb._param_keys, ref_params):
            p.grad = bb._grad.view(k).detach().clone()
        opt.step(); ref.step()
        b = opt._built
        now = (b["lr"].data_ptr(), b["wd"].data_ptr(), b["ranges"].data_ptr(), b["mom"].data_ptr())
        if step == 0:
            ptrs = now
            assert float(b["mom"].abs().sum()) > 0
            opt.set_lr(net.parameter_groups(0.5 * cfg.LR, cfg.WEIGHT_DECAY))        # schedule step
            for gr in ref.param_groups:
                gr["lr"] *= 0.5
            assert opt._built is b and float(b["mom"].abs().sum()) > 0, "set_lr dropped the momentum buffer"
        assert now == ptrs, "set_lr moved tensors a captured graph points at"
    got = torch.cat([p.detach().reshape(-1) for p in bb.parameters()])
    want = torch.cat([p.detach().reshape(-1) for p in ref_params])
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-6


def test_tail_and_loss_kernel_forms_are_bit_identical_on_the_emulation():
    """tests/test_tail_loss_variants_gpu.py on the host emulation (real kernel source, thin geometries): staged vs direct
    up-sampling -- two subprocesses, digests of every output tensor must agree"""
    m = importlib.import_module("test_tail_loss_variants_gpu")
    m.check_variants({"SACB_PROBE_EMUL": "1"})
