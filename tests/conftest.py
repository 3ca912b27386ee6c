"""pytest configuration: the ``gpu`` marker and golden-fixture helpers."""
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


def load_golden(group):
    """tests/golden/<group>.npz -> {case: {key: ndarray}} (written by oracle/gen_golden.py)."""
    z = np.load(os.path.join(GOLDEN, group + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        case, key = k.split("::", 1)
        out.setdefault(case, {})[key] = z[k]
    return out


def cfg_of(rec):
    return {k[4:]: (v.item() if v.shape == () else v) for k, v in rec.items() if k.startswith("cfg_")}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max() if b.size else 0.0, 1e-30)
    return (np.abs(a - b).max() if b.size else 0.0) / denom


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(group):
        if group not in cache:
            cache[group] = load_golden(group)
        return cache[group]
    return get
