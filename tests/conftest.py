import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def rel_err(a, b):
    """Parity metric of SURVEY.md section 4: max|a-b| / max|b| per tensor."""
    import numpy as np
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    assert a.size == b.size, (a.size, b.size)
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    return d / max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-30)


@pytest.fixture(scope="session")
def cenn():
    """The product package with a live cenn_state on cuda:0 (GPU tests only)."""
    import video_filler_b200.tensor as T
    T.state(0)
    return T
