"""End-to-end: exec/icp_test.cpp replayed on the bundled 121-frame dataset with every numeric step
on the GPU (host buffers through the C-ABI), against the oracle replay and the reference's output/."""
import numpy as np
import pytest

import backends
import replay

pytestmark = pytest.mark.gpu


def test_dataset_replay(dataset):
    gpu = replay.run_icp_test(dataset, backends.GpuBackend())
    cpu = replay.run_icp_test(dataset, backends.OracleBackend())
    # same map: 490 landmarks with identical ids in identical order (matching is bit-exact)
    assert len(gpu["world"].xyz) == 490
    assert np.array_equal(gpu["world"].id_real, cpu["world"].id_real)
    assert np.array_equal(gpu["inliers"][:, 1], cpu["inliers"][:, 1])
    # iteration counts are numerics-sensitive (1e-5 stop at float noise): poses are compared instead
    dpos = np.abs(gpu["poses"] - cpu["poses"]).max()
    assert dpos < 5e-3, dpos
    ev = replay.evaluate(dataset, gpu)
    g = dataset
    dxy = np.linalg.norm(ev["traj"][:, 1:3] - g["golden_traj"][:, 1:3], axis=1).max()
    dth = np.abs(ev["traj"][:, 3] - g["golden_traj"][:, 3]).max()
    assert np.array_equal(ev["world_points"][:, 0], g["golden_world_points"][:, 0])
    assert dxy <= 0.006 * 41.4, dxy  # <= 0.6 % of the extent, <= 0.012 rad (RANSAC 5-point first pose)
    assert dth <= 0.012, dth
