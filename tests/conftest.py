import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box through gpurun)")


def pytest_collection_modifyitems(config, items):
    """Order under `pytest -x`: the multi-GPU file last (a failure there must not hide the single-GPU parity results)."""
    def rank(item):
        name = item.nodeid
        if "test_gpu_z_multi" in name:
            return 3
        return 0
    items.sort(key=rank)  # stable: the original order is kept inside each group


@pytest.fixture(scope="session")
def port():
    from oracle import api
    return api.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import api
    r = api.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (needs /root/reference; `make -C oracle`)")
    return r


@pytest.fixture(scope="session")
def ref_serial():
    from oracle import api
    r = api.ref(serial=True)
    if r is None:
        pytest.skip("oracle/_ref not built")
    return r


@pytest.fixture(scope="session")
def golden_logs():
    with open(os.path.join(GOLDEN, "ref_logs.json")) as f:
        return json.load(f)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_l2(a, b):
    """relative L2 error of a against reference b (the north_star parity measure)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.sqrt((b * b).sum())
    num = np.sqrt(((a - b) ** 2).sum())
    return num / den if den > 0 else num


def sine_rhs(n, m=None):
    m = n if m is None else m
    x, y = np.arange(n) / n, np.arange(m) / m
    return -2 * np.pi ** 2 * np.outer(np.sin(np.pi * x), np.sin(np.pi * y))
