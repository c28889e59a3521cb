import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def analytic():
    return load_golden("analytic.npz")


@pytest.fixture(scope="session")
def obs_spaces():
    return load_golden("obs_spaces.npz")


@pytest.fixture(scope="session")
def hopf():
    return load_golden("hopf.npz")


ROLLOUTS = sorted(f[len("rollout_"):-4] for f in os.listdir(GOLDEN) if f.startswith("rollout_"))
