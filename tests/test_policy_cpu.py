"""Host-side policies that need no GPU: the precision policy of lib.py and the invariant that the GEMM / exchange kernels of
the default path are instruction-identical to the last build that passed ``pytest -m gpu`` on a B200."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VERIFIED_COMMIT = "289435f"          # profiles/pytest_gpu_r1q.log, profiles/bench_r1q.json


def test_precision_policy(monkeypatch):
    from da_sac_b200 import lib as L
    table = {("parity", "fwd"): 0, ("parity", "bwd"): 0, ("fast_bwd", "fwd"): 0, ("fast_bwd", "bwd"): 1, ("fast", "fwd"): 1, ("fast", "bwd"): 1}
    for (policy, phase), want in table.items():
        monkeypatch.setattr(L, "PRECISION", policy)
        L.set_phase(phase)
        assert L._precision() == want, (policy, phase)
    L.set_phase("fwd")
    assert os.environ.get("SACB_PRECISION", "parity") == "parity", "the test suite runs in the parity mode"


def test_default_kernels_are_identical_to_the_gpu_verified_build():
    """switch-gated additions to sacb_gemm.cu / sacb_p2p.cu must be NEW template instantiations only (profiles/sass_identity.py)"""
    if shutil.which("nvcc") is None or shutil.which("cuobjdump") is None or shutil.which("git") is None:
        pytest.skip("needs nvcc, cuobjdump and git")
    have = subprocess.run(["git", "-C", ROOT, "cat-file", "-e", VERIFIED_COMMIT + "^{commit}"], capture_output=True)
    if have.returncode != 0:
        pytest.skip("commit %s is not in this checkout" % VERIFIED_COMMIT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sass_identity.py"), VERIFIED_COMMIT],
                       capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all default kernels identical" in r.stdout
