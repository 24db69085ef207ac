"""Shared test plumbing.

Markers:  @pytest.mark.gpu — needs a CUDA device (run on the B200 box with `-m gpu`);
everything else runs on CPU (`-m "not gpu"`).
The oracle (oracle/) is imported here and in the tests only — it is the checker, never the thing under test's
implementation.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


def have_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:      # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    if have_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "g*.npz")))


@pytest.fixture(scope="session")
def golden():
    """name -> (Scene, npz dict) for every fixture under tests/golden/."""
    from gipuma_b200.golden import scene_from_arrays
    out = {}
    for name in golden_names():
        z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
        out[name] = (scene_from_arrays(name, z), z)
    return out


def bits_equal(a, b) -> int:
    """Number of elements whose bit patterns differ (NaN == NaN)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape
    neq = a.view(np.uint32) != b.view(np.uint32)
    neq &= ~(np.isnan(a) & np.isnan(b))
    return int(neq.sum())
