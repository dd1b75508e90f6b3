import os
import sys

# The host emulation (tests/cpu_emul) runs its own pool of OS threads next to torch's intra-op threads; OpenMP workers that
# spin after every parallel region then fight that pool for the cores and the suite's run time becomes erratic.  Must be set
# before torch (libgomp) starts its first parallel region; a no-op for the GPU tests.
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
os.environ.setdefault("GOMP_SPINCOUNT", "0")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "sac_resnet101_tiny.npz"), allow_pickle=False)
