import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nn-fac_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture
def fp32_path():
    """Small float32 problems compute in float64 in "auto" mode (nn_fac/config.py: they are launch-bound, the reference's own
    arithmetic is free there).  Tests that target the float32 / tensor-core kernels at small sizes switch that off."""
    import nn_fac.config as config
    old = config.small_problem_elements
    config.small_problem_elements = 0
    yield
    config.small_problem_elements = old
