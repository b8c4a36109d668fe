import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build everything once per session (same entry point the driver uses)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def matr33():
    with open(os.path.join(ROOT, "tests", "golden", "matr33.json")) as f:
        g = json.load(f)
    g["rows"] = np.array(g["rows"], np.int32)
    g["cols"] = np.array(g["cols"], np.int32)
    g["vals"] = np.array(g["vals"], np.float64)
    g["b"] = np.array(g["b"], np.float64)
    g["x_golden"] = np.array(g["x_golden"])
    g["x_direct"] = np.array(g["x_direct"])
    return g
