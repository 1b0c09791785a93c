import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return lambda name: np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def engine():
    from pose2sim_b200 import ops
    return ops.get_engine(0)


def tri_cases(npz, prefix_fmt, n):
    """Yield (name, P, x, y, w, thr, min_cams, Q, err, nexcl, mask) from a golden triangulation file."""
    for i in range(n):
        p = prefix_fmt.format(i)
        thr, mc = npz[p + "params"]
        yield (p, npz[p + "P"], npz[p + "x"], npz[p + "y"], npz[p + "w"], float(thr), int(mc),
               npz[p + "Q"], npz[p + "err"], npz[p + "nexcl"], npz[p + "mask"])
