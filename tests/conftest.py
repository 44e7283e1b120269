import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def kitti_pair():
    """Two consecutive synthetic KITTI-shaped frames (1241x376), seed 0."""
    from monocular_visual_odometry_va4mr_b200 import synth
    return synth.render_sequence("kitti", 2, seed=0)


@pytest.fixture(scope="session")
def small_pair():
    """Small synthetic pair (320x240) used by the golden fixtures."""
    d = np.load(os.path.join(GOLDEN, "klt_small.npz"))
    return d


@pytest.fixture(scope="session")
def ctx():
    from monocular_visual_odometry_va4mr_b200 import _lib
    return _lib.default_context(0)
