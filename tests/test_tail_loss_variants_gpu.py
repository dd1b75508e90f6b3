"""The teacher tail's probability kernel and the student-loss forward kernel exist in two forms -- low-resolution logits read
from global memory, or staged in shared memory (the default; ``SACB_UP_STAGED=0`` selects the first).  The forms must agree BIT
FOR BIT (same expression per class): the labels of the direct form were pinned to the real reference's golden in round 1
(tests/test_step_gpu.py), identity carries the pin over, on ragged geometries too (rows that are not a multiple of the
256-pixel block, blocks that span several rows, staging that does not fit and falls back).  The loss backward is probed along
the way (digest of d logits for the self-training and the plain cross-entropy form).

The switch is read once per process, hence one subprocess per variant; each prints a digest of every output tensor."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r'''
import contextlib, ctypes as C, hashlib, os, sys
import torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
from da_sac_b200 import lib as L, synth
from da_sac_b200.models import get_model
ctx = contextlib.nullcontext()
if os.environ.get("SACB_PROBE_EMUL") == "1":           # the GPU-less container: same probe on the host emulation (tests/emul_harness.py)
    import emul_harness as E
    ctx = E.emulated_gpu()
ctx.__enter__()
dev = torch.device("cuda")
GEOMS = [(2, 3, (512, 512)), (1, 2, (640, 640)), (2, 2, (97, 131)), (1, 3, (70, 203)), (1, 2, (128, 1024)), (1, 2, (64, 1000))]
if os.environ.get("SACB_PROBE_EMUL") == "1":
    GEOMS = [(1, 2, (64, 512)), (1, 2, (40, 640)), (2, 2, (97, 131)), (1, 3, (70, 203)), (1, 2, (24, 1024)), (1, 2, (24, 1000))]
# (.., 1000): 256-pixel blocks straddle rows of 126 low-resolution columns -> three staged rows do not fit -> the direct form inside the staged build
def low_res(n):
    n = (n - 1) // 2 + 1; n = -(-(n - 1) // 2) + 1
    return (n - 1) // 2 + 1
def sha(t): return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]
cfg = synth.ModelCfg()
m = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
m.cuda().train()
for G, K, (H, W) in GEOMS:
    BT = G * K
    g = torch.Generator().manual_seed(H * 7 + W)
    x, y, x2, A, Ai = [t.cuda() for t in synth.make_target_batch(G, K, (H, W), seed=1)]
    h, w = low_res(H), low_res(W)
    t_logits = (torch.randn(BT, 19, h, w, generator=g) * 3).cuda()
    s_logits = (torch.randn(BT, 19, h, w, generator=g) * 3).cuda()
    m.running_conf.fill_(0.02)
    ws = m._tail(t_logits, y, A, Ai, K)
    torch.cuda.synchronize()
    tag = "%%dx%%dx%%d" %% (BT, H, W)
    for k in ("probs", "labels", "conf", "conf_mean", "thresholds"):
        print("sha", tag, "tail." + k, sha(ws[k]))
    print("sha", tag, "tail.running_conf", sha(m.running_conf))
    losses = torch.zeros(2, device=dev); scratch = torch.zeros(2, dtype=torch.float64, device=dev)
    grows = torch.empty(BT * 19 * H * w, device=dev)
    yy = y.clone(); yy[yy == -1] = 255
    for name, labels, scale in (("self_ce", ws["labels"], 5.0), ("loss_ce", None, 1.0)):
        dl = torch.zeros_like(s_logits)
        d = L.Loss(C.sizeof(L.Loss), BT, 19, h, w, H, W, L.ptr(s_logits), L.ptr(yy), L.ptr(labels) if labels is not None else None,
                   L.ptr(ws["conf_mean"]), L.ptr(m.running_conf), 3.0, L.ptr(losses), L.ptr(scratch), scale, L.ptr(dl), None, L.ptr(grows))
        L.check(L.lib().sacb_student_loss_fwd(C.byref(d), L.stream()), "fwd")
        L.check(L.lib().sacb_student_loss_bwd(C.byref(d), L.stream()), "bwd")
        torch.cuda.synchronize()
        print("sha", tag, "dlogits." + name, sha(dl))
        print("val", tag, "losses." + name, "%%.9e %%.9e" %% (losses[0].item(), losses[1].item()))
print("launches", L.launch_count())
ctx.__exit__(None, None, None)
'''


def run(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", PROBE % (ROOT, ROOT)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    sha = [ln for ln in r.stdout.splitlines() if ln.startswith("sha ")]
    val = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("val ")]
    assert len(sha) == 6 * 8 and len(val) == 6 * 2 and "launches" in r.stdout
    return sha, val


def check_variants(common):
    base_sha, base_val = run(dict(common, SACB_UP_STAGED="0"))
    for extra in ({},):
        sha, val = run(dict(common, **extra))
        diff = [(a, b) for a, b in zip(base_sha, sha) if a != b]
        assert not diff, "variant %r differs from the direct forms: %s" % (extra, diff[:4])
        for a, b in zip(base_val, val):                      # the two loss sums are fp64 atomics across blocks: order-free to 1e-12
            assert a[:3] == b[:3]
            for u, v in zip(a[3:], b[3:]):
                assert abs(float(u) - float(v)) <= 1e-6 * max(abs(float(u)), 1e-6), (a, b)


def test_staged_form_is_bit_identical_to_the_direct_form():
    check_variants({})
