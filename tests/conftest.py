import glob
import importlib
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


def pkg(sub=None):
    """The product package (its directory name starts with a digit, so it is imported by string)."""
    return importlib.import_module("3danimals_b200" + ("." + sub if sub else ""))


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_files(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def elem_err(a, b, rtol=1e-4, floor=1e-6):
    """Per-ELEMENT check |a - b| <= rtol * |b| + floor * max|b| (north_star: "1e-4 relative on float buffers"; the floor keeps
    entries that are zero up to rounding from demanding infinite relative accuracy).  -> (ok, worst ratio, index of the worst
    element): ratio = |a - b| / (rtol |b| + floor max|b|), so <= 1 passes."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    bound = rtol * np.abs(b) + floor * max(float(np.abs(b).max()), 1e-30)
    ratio = np.abs(a - b) / bound
    i = int(np.argmax(ratio)) if ratio.size else 0
    worst = float(ratio.reshape(-1)[i]) if ratio.size else 0.0
    return worst <= 1.0, worst, np.unravel_index(i, ratio.shape) if ratio.size else ()


def assert_elem(a, b, rtol=1e-4, floor=1e-6, what=""):
    ok, worst, idx = elem_err(a, b, rtol, floor)
    assert ok, "%s: worst element %s is %.2f x the per-element bound (rtol %g, floor %g of max)" % (what, idx, worst, rtol, floor)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
