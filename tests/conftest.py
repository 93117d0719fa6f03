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
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fixtures():
    return json.load(open(os.path.join(GOLDEN, "fixtures.json")))


@pytest.fixture(scope="session")
def bag_clouds():
    """bag index -> (N,8) float32 records in the PointCloud2 layout of the bundled bags
    (x@0 y@4 z@8 pad intensity@16 ring(u16)@20 pad), rebuilt from tests/golden/bag_clouds.npz."""
    z = np.load(os.path.join(GOLDEN, "bag_clouds.npz"))
    out = {}
    for b in range(1, 7):
        xyzi, ring = z[f"xyzi_{b}"], z[f"ring_{b}"]
        rec = np.zeros((xyzi.shape[0], 8), np.float32)
        rec[:, 0:3] = xyzi[:, 0:3]
        rec[:, 4] = xyzi[:, 3]
        rec.view(np.uint16).reshape(-1, 16)[:, 10] = ring
        out[b] = rec
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def ensure_product_built():
    from lidar_camera_calibration_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "lidar_camera_calibration_b200", "csrc"), "-s"])
    return _lib.load()


@pytest.fixture(scope="session")
def product_lib():
    return ensure_product_built()
