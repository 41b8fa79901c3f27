import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def dataset():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "dataset.npz")))


@pytest.fixture(scope="session")
def cv2fx():
    return np.load(os.path.join(ROOT, "tests", "golden", "cv2_fixtures.npz"))


@pytest.fixture(scope="session")
def cv2tri():
    """cv2.triangulatePoints on 24 adversarial two-view problems (oracle/gen_golden_tri.py)"""
    return np.load(os.path.join(ROOT, "tests", "golden", "cv2_triangulate.npz"))


@pytest.fixture(scope="session")
def cv2pose():
    """cv2.findEssentialMat(RANSAC) + recoverPose on 20 synthetic two-view problems (oracle/gen_golden_pose.py)"""
    return np.load(os.path.join(ROOT, "tests", "golden", "cv2_recoverpose.npz"))


@pytest.fixture(scope="session")
def world_gt():
    """data/world.dat + the camera-in-robot transform of data/camera.dat (oracle/gen_golden_world.py)"""
    return np.load(os.path.join(ROOT, "tests", "golden", "world_gt.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle
