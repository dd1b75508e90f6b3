"""CPU dry run of the host-side layer schedules: every C-ABI call is replaced by a recorder that returns success, tensors live
on the CPU.  No arithmetic is checked here (that is what the GPU parity tests do) -- this catches what only shows up when the
Python schedule actually executes: wrong tags in the scratch-plane pools, pool exhaustion, missing keys, argument-count /
ctypes conversion errors, autograd wiring.  Covers the frozen-BN ResNet-101 engine (the verified SAC path, as a sanity check of
the harness) and the training-BN engine of the ABN baseline (engine_abn.py)."""
import collections
import ctypes as C

import pytest
import torch


class FakeLib(object):
    def __init__(self):
        self.calls = collections.Counter()

    def __getattr__(self, name):
        if not name.startswith("sacb_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls[name] += 1
            for a in args:                                   # every argument must be something ctypes can pass
                assert a is None or isinstance(a, (int, C._SimpleCData, C.Array, type(C.byref(C.c_int(0))))), (name, type(a))
            if name == "sacb_bn_moments_partial_elems":
                M, Cn = args
                return ((M.value + 511) // 512) * 2 * Cn
            if name == "sacb_conv_wgrad_splits":
                return 3
            if name == "sacb_prep_item_blocks":
                return 2
            if name == "sacb_abi_version":
                return 4
            if name == "sacb_launch_count":
                return sum(self.calls.values())
            if name in ("sacb_tail_part_sums_elems", "sacb_tail_probs_elems", "sacb_tail_pooled_elems"):
                return 1024
            return 0
        return fn


@pytest.fixture
def fake(monkeypatch):
    from da_sac_b200 import lib as L
    f = FakeLib()
    monkeypatch.setattr(L, "lib", lambda: f)
    monkeypatch.setattr(L, "stream", lambda: C.c_void_p(0))
    monkeypatch.setattr(L, "ptr", lambda t: None if t is None else C.c_void_p(t.data_ptr()))
    monkeypatch.setattr(L, "dptr", lambda t: None if t is None else t.data_ptr())
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)     # no CUDA runtime in the CPU container
    return f


def _net(baseline):
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model

    class Cfg(synth.ModelCfg):
        BASELINE = baseline
    return get_model(Cfg(), 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none")), Cfg()


def test_frozen_bn_engine_schedule_runs(fake):
    from da_sac_b200 import engine as E
    net, _ = _net(False)
    bb = net.backbone
    bb.train()
    x = torch.randn(2, 3, 65, 65)
    logits = bb.logits(x)                                    # student forward through _BackboneFn
    assert logits.shape == (2, 19, 9, 9) and logits.requires_grad
    logits.sum().backward()
    assert all(p.grad is not None for p in bb.parameters())
    eng = bb.engine(2, 65, 65)
    assert type(eng) is E.ResNet101Engine
    assert fake.calls["sacb_conv_gemm"] > 200 and fake.calls["sacb_conv_wgrad"] > 100
    assert fake.calls["sacb_bn_moments"] == 0                # frozen BN: folded into the GEMM epilogue


def test_training_bn_engine_schedule_runs(fake):
    from da_sac_b200 import engine_abn, synth
    net, cfg = _net(True)
    net.train()
    bb = net.backbone
    assert bb._bn_training()
    xs, ys = synth.make_source_batch(2, (65, 65), seed=0)
    # source step: forward with autograd, fused CE loss, backward
    losses, outs = net(xs, ys)
    eng = bb.engine(2, 65, 65)
    assert type(eng) is engine_abn.ResNet101TrainBNEngine and not bb._wp.fold_bn
    n_units = len(eng.units)
    assert n_units == 104                                    # stem + 33 x (conv1, conv2, conv3) + 4 downsample convs
    assert fake.calls["sacb_bn_moments"] == n_units and fake.calls["sacb_bn_apply"] == n_units
    assert fake.calls["sacb_bn_train_finalize"] == n_units
    assert int(bb.model.layer3[5].bn2.num_batches_tracked) == 1
    losses["loss_ce"].mean().backward()
    assert fake.calls["sacb_bn_moments"] == 2 * n_units and fake.calls["sacb_bn_bwd_apply"] == n_units
    assert fake.calls["sacb_bn_bwd_finalize"] == n_units
    assert fake.calls["sacb_student_loss_fwd"] == 1 and fake.calls["sacb_student_loss_bwd"] == 1
    assert all(p.grad is not None for p in bb.parameters())
    assert set(outs.keys()) >= {"logits", "logits_up"}
    # ABN target pass: no-grad forward in train mode (scratch planes only), statistics kernels still run
    before = fake.calls["sacb_bn_train_finalize"]
    with torch.no_grad():
        losses_t, _ = net(xs, ys)
    assert fake.calls["sacb_bn_train_finalize"] == before + n_units
    assert int(bb.model.layer3[5].bn2.num_batches_tracked) == 2
    # eval: frozen-BN engine, folded planes again
    net.eval()
    calls = fake.calls["sacb_bn_moments"]
    logits, up = net(xs)
    assert fake.calls["sacb_bn_moments"] == calls and bb._wp.fold_bn
    assert logits.shape == (2, 19, 9, 9)


@pytest.mark.parametrize("arch,n_units", [("vgg16", 13), ("fcn", 15)])
def test_training_bn_schedules_of_the_vgg_backbones_run(fake, arch, n_units):
    """the VGG engines get training-mode BN through their _unit / _wgrad building blocks (engine_abn._VGGTrainBN)"""
    from da_sac_b200 import engine_abn, synth
    from da_sac_b200.models import get_model
    Base = {"vgg16": synth.ModelCfgVGG16, "fcn": synth.ModelCfgFCN}[arch]

    class Cfg(Base):
        BASELINE = True
    net = get_model(Cfg(), 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    net.train()
    bb = net.backbone
    xs, ys = synth.make_source_batch(2, (64, 64), seed=0)
    losses, outs = net(xs, ys)
    eng = bb.engine(2, 64, 64)
    assert type(eng) is {"vgg16": engine_abn.VGG16TrainBNEngine, "fcn": engine_abn.FCN8sTrainBNEngine}[arch]
    assert len(eng.units) == n_units and not bb._wp.fold_bn
    assert fake.calls["sacb_bn_moments"] == n_units == fake.calls["sacb_bn_apply"] == fake.calls["sacb_bn_train_finalize"]
    losses["loss_ce"].mean().backward()
    assert fake.calls["sacb_bn_moments"] == 2 * n_units and fake.calls["sacb_bn_bwd_apply"] == n_units
    assert all(p.grad is not None for p in bb.parameters())
    with torch.no_grad():
        net(xs, ys)
    assert fake.calls["sacb_bn_train_finalize"] == 2 * n_units
    net.eval()
    before = fake.calls["sacb_bn_moments"]
    net(xs)
    assert fake.calls["sacb_bn_moments"] == before and bb._wp.fold_bn


@pytest.mark.parametrize("arch", ["resnet101", "vgg16", "fcn"])
def test_full_sac_target_step_schedule_runs(fake, arch):
    """SAC.forward (EMA / teacher forward / tail / student forward / fused loss) + backward for every backbone: the verified
    default path, exercised here so that host-side edits to it are caught without a GPU"""
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    cfg = {"resnet101": synth.ModelCfg, "vgg16": synth.ModelCfgVGG16, "fcn": synth.ModelCfgFCN}[arch]()
    net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    net.train()
    assert not net.backbone._bn_training() and not net.slow_net._bn_training()
    K, HW = 2, (64, 64)
    x, y, x2, A, Ai = synth.make_target_batch(1, K, HW, seed=0)
    for step in range(2):
        losses, outs = net(x.clone(), y.clone(), x2.clone(), A, Ai, use_teacher=True, update_teacher=(step == 0), T=K)
        assert set(losses) >= {"loss_ce", "self_ce", "teacher_diff"}
        net.zero_grad()
        (cfg.LR_TARGET * losses["self_ce"].mean()).backward()
        assert all(p.grad is not None for p in net.backbone.parameters())
    assert fake.calls["sacb_teacher_tail"] == 2 and fake.calls["sacb_student_loss_bwd"] == 2
    assert fake.calls["sacb_bn_moments"] == 0 and net.backbone._wp.fold_bn
    # source pass of the joint recipe (train.py:119-138 with BASELINE = False): loss_ce backward
    losses, _ = net(x.clone(), torch.zeros_like(y))
    losses["loss_ce"].mean().backward()
