import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    return load


def force_tolerance_ok(f, f_ref, rel=1e-5, abs_=1e-6):
    """north_star: forces within 1e-5 relative plus 1e-6 eV/A absolute."""
    return np.all(np.abs(f - f_ref) <= rel * np.abs(f_ref) + abs_)


def max_force_violation(f, f_ref, rel=1e-5, abs_=1e-6):
    return float(np.max(np.abs(f - f_ref) / (rel * np.abs(f_ref) + abs_)))
