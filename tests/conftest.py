import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def sort_agents(a: dict) -> dict:
    """Canonical order for comparing populations as sets: by agent id."""
    o = np.argsort(a["id"], kind="stable")
    return {k: v[o] for k, v in a.items()}


@pytest.fixture(scope="session")
def small_world():
    from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude
    nbr, xyz = make_ico_grid(15)  # 2562 cells
    alt = synthetic_altitude(xyz, seed=3)
    return nbr, xyz, alt
