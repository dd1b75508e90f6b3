"""GPU parity of the fractional-group tail (a view-group spread over two ranks) against golden vectors the REAL
reference produced on two gloo ranks.  One GPU plays both ranks: the sub-group all-reduce of the reference-frame partial
sums is emulated by adding the two ranks' ``pooled`` buffers (the collective itself is NCCL plumbing; its host side is
covered on gloo in tests/test_fractional_cpu.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
K, HW, WORLD = 4, (96, 96), 2


def test_fractional_tail_labels_bit_exact_on_golden_teacher_logits(monkeypatch):
    from da_sac_b200 import synth
    from da_sac_b200.models import get_model
    g = np.load(os.path.join(HERE, "golden", "sac_fractional_w2.npz"))
    cfg = synth.ModelCfg()
    batch = synth.make_target_batch(1, K, HW, seed=3)
    per = K // WORLD
    nets = []
    for r in range(WORLD):
        m = get_model(cfg, r, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
        m.cuda().train()
        m.running_conf.fill_(cfg.THRESHOLD_BETA)
        nets.append(m)
    SAC = type(nets[0])
    ins = []
    for r in range(WORLD):
        _, y, _, A, Ai = [t[r * per:(r + 1) * per].cuda() for t in batch]
        ins.append((torch.from_numpy(g["r%d_teacher_logits" % r]).cuda(), y, A, Ai))
    # pass 1: every rank's un-normalised partial sums, snapshotted at the exchange point (eval mode: running_conf untouched)
    snaps = {}

    def snap_exchange(self, pooled, B, T):
        assert (B, T) == (per, K)
        snaps[self.rank] = pooled.clone()
    monkeypatch.setattr(SAC, "_exchange_partial_sums", snap_exchange, raising=True)
    for m, (tl, y, A, Ai) in zip(nets, ins):
        m.training = False
        m._tail(tl, y, A, Ai, K)
    torch.cuda.synchronize()
    total = snaps[0] + snaps[1]                               # what the sub-group all-reduce delivers to both ranks
    monkeypatch.setattr(SAC, "_exchange_partial_sums", lambda self, pooled, B, T: pooled.copy_(total), raising=True)
    # pass 2: the real tail of every rank
    for r, (m, (tl, y, A, Ai)) in enumerate(zip(nets, ins)):
        m.training = True
        m.running_conf.fill_(cfg.THRESHOLD_BETA)
        ws = m._tail(tl, y, A, Ai, K)
        torch.cuda.synchronize()
        rc = torch.from_numpy(g["r%d_running_conf" % r])
        assert torch.allclose(m.running_conf.cpu(), rc, rtol=1e-5, atol=1e-8)
        conf, gconf = ws["conf"].cpu(), torch.from_numpy(g["r%d_teacher_conf" % r])
        assert (conf - gconf).abs().max() < 2e-5, (conf - gconf).abs().max()
        lab, glab = ws["labels"].cpu(), torch.from_numpy(g["r%d_teacher_labels" % r])
        amb = torch.from_numpy(g["r%d_ambiguous" % r])
        mism = lab != glab
        print("rank", r, "label mismatches", int(mism.sum()), "ambiguous", int(amb.sum()))
        assert int((mism & ~amb).sum()) == 0
        # refined probabilities through the lazy diagnostic path (phase 1 + exchange + phase 2 with `refined`)
        refined = torch.empty(per, 19, HW[0], HW[1], device="cuda")
        m.training = False
        m._tail(tl, y, A, Ai, K, refined=refined)
        torch.cuda.synchronize()
        gsub = torch.from_numpy(g["r%d_teacher_refined_sub" % r])
        assert (refined[:, :, ::3, ::3].cpu() - gsub).abs().max() < 2e-5


def test_whole_group_tail_unchanged_by_phase_split():
    """phase 1 + (no-op exchange) + phase 2 on a WHOLE group reproduces the single-call tail bit for bit"""
    import ctypes as C
    from da_sac_b200 import lib as L, synth
    from da_sac_b200.models import get_model
    cfg = synth.ModelCfg()
    m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
    m.cuda().eval()
    m.running_conf.fill_(0.05)
    Kk, hw = 3, (64, 80)
    _, y, _, A, Ai = [t.cuda() for t in synth.make_target_batch(2, Kk, hw, seed=7)]
    torch.manual_seed(0)
    tl = (torch.randn(2 * Kk, 19, 9, 11) * 4).cuda()
    ws = m._tail(tl, y, A, Ai, Kk)
    torch.cuda.synchronize()
    lab0, conf0 = ws["labels"].clone(), ws["conf"].clone()

    def desc(phase):
        return L.Tail(C.sizeof(L.Tail), 2 * Kk, Kk, 19, 9, 11, hw[0], hw[1], L.ptr(tl), L.ptr(y), L.ptr(A.contiguous()),
                      L.ptr(Ai.contiguous()), L.ptr(m.running_conf), 0, 1 if cfg.CONF_DISCOUNT else 0, cfg.THRESHOLD_BETA,
                      cfg.STAT_MOMENTUM, cfg.RUN_CONF_UPPER, cfg.RUN_CONF_LOWER, L.ptr(ws["probs"]), L.ptr(ws["pooled"]),
                      L.ptr(ws["part_sums"]), L.ptr(ws["peaks"]), L.ptr(ws["conf"]), L.ptr(ws["idx"]), L.ptr(ws["labels"]),
                      L.ptr(ws["conf_mean"]), L.ptr(ws["thresholds"]), None, phase)
    ws["labels"].zero_(); ws["conf"].zero_()
    L.check(L.lib().sacb_teacher_tail(C.byref(desc(1)), L.stream()), "phase 1")
    L.check(L.lib().sacb_teacher_tail(C.byref(desc(2)), L.stream()), "phase 2")
    torch.cuda.synchronize()
    assert torch.equal(ws["labels"], lab0) and torch.equal(ws["conf"], conf0)
