import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "reference_runs.npz"))


@pytest.fixture(scope="session")
def frame0_xyz():
    return np.load(os.path.join(GOLDEN, "frame0_xyz.npy"))


@pytest.fixture(scope="session")
def frame0_h5_xyz():
    """enspara/test/data/frame0.h5 '/coordinates' (decoded by enspara_b200/util/h5min.py,
    scripts/make_golden.py): the fixture of the reference's PAM goldens."""
    return np.load(os.path.join(GOLDEN, "frame0_h5_xyz.npy"))
