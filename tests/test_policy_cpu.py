"""Host-side policies that need no GPU: the precision policy of lib.py.  (Round 1 also pinned the default kernels' SASS to the
last GPU-verified build, because kernel edits could not be run any more; in round 2 kernels change and are re-verified on the
B200 instead -- ``profiles/sass_identity.py`` stays as a tool.)"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_precision_policy(monkeypatch):
    from da_sac_b200 import lib as L
    table = {("parity", "fwd"): 0, ("parity", "bwd"): 0, ("fast_bwd", "fwd"): 0, ("fast_bwd", "bwd"): 1, ("fast", "fwd"): 1, ("fast", "bwd"): 1}
    for (policy, phase), want in table.items():
        monkeypatch.setattr(L, "PRECISION", policy)
        L.set_phase(phase)
        assert L._precision() == want, (policy, phase)
    L.set_phase("fwd")
    assert os.environ.get("SACB_PRECISION", "parity") == "parity", "the test suite runs in the parity mode"
