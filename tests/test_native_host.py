"""The C++ host mirror of the reference interface (02-visualodometry_b200/host): the native replay of
exec/icp_test.cpp on the bundled dataset, run as the reference is run (a process reading meas-*.dat and
writing output/*.txt)."""
import os
import subprocess

import numpy as np
import pytest

import backends
import dataset_io
import replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "02-visualodometry_b200", "host", "icp_test_native")


VO_BIN = os.path.join(ROOT, "02-visualodometry_b200", "host", "vo_native")


def test_native_binary_is_built():
    assert os.path.exists(BIN), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    assert os.path.exists(VO_BIN)


@pytest.mark.gpu
def test_native_vo_driver(dataset, tmp_path):
    """exec/vo.cpp's flow (Cam::initOneRound/oneRound: threshold 1000, five rounds): the C++ mirror, the Python
    replay through the C-ABI and the oracle replay agree; free oneRound()'s parameters (threshold 100, outliers
    kept) are covered by tests/test_gpu_picp.py."""
    prefix = dataset_io.write_meas_files(dataset, str(tmp_path / "data"), n_meas=120)
    out = tmp_path / "vo_poses.txt"
    r = subprocess.run([VO_BIN, prefix, str(out), "120"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    native = np.loadtxt(out).reshape(-1, 3, 4)
    assert native.shape[0] == 120
    gpu = replay.run_vo(dataset, backends.GpuBackend())
    cpu = replay.run_vo(dataset, backends.OracleBackend())
    assert np.abs(native - gpu["poses"]).max() <= 1e-5          # same calls, printed with 9 digits
    assert len(gpu["world"].xyz) == len(cpu["world"].xyz)
    assert np.array_equal(gpu["world"].id_real, cpu["world"].id_real)
    assert np.array_equal(gpu["inliers"][:, 1], cpu["inliers"][:, 1])
    assert np.abs(gpu["poses"] - cpu["poses"]).max() <= 5e-3
    assert "Number of duplicate world points" in r.stdout


def test_native_refuses_to_run_without_gpu(dataset, tmp_path):
    """no CPU fallback: without a device the process must fail loudly, not produce output files"""
    vo = backends.product()
    if vo.device_count() > 0:
        pytest.skip("a GPU is visible")
    prefix = dataset_io.write_meas_files(dataset, str(tmp_path / "data"), n_meas=3)
    out = tmp_path / "output"
    out.mkdir()
    r = subprocess.run([BIN, prefix, str(out), "3"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "no usable CUDA device" in (r.stderr + r.stdout)
    assert not (out / "estimated_trajectory.txt").exists() or os.path.getsize(out / "estimated_trajectory.txt") == 0


@pytest.mark.gpu
def test_native_icp_test_reproduces_output(dataset, tmp_path):
    prefix = dataset_io.write_meas_files(dataset, str(tmp_path / "data"))
    out = tmp_path / "output"
    out.mkdir()
    r = subprocess.run([BIN, prefix, str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = dataset_io.read_outputs(str(out))
    g = dataset
    # the four files have the reference's shapes; the map is the reference's 490 landmarks
    assert got["traj"].shape == (121, 4) and got["errors"].shape == (121, 3)
    assert got["world_points"].shape == (490, 4)
    assert np.array_equal(got["world_points"][:, 0], g["golden_world_points"][:, 0])
    # the summary lines the reference prints are still there (first frame pair: 115 matches, all correct)
    assert "Matches: Out of 115 possible matches, found 115, of which 115 are correct" in r.stdout
    dxy = np.linalg.norm(got["traj"][:, 1:3] - g["golden_traj"][:, 1:3], axis=1).max()
    dth = np.abs(got["traj"][:, 3] - g["golden_traj"][:, 3]).max()
    assert dxy <= 0.01 * 41.4, dxy      # 8-point initial E, see DESIGN.md section 2
    assert dth <= 0.015, dth
    assert np.abs(got["errors"][:, 1] - g["golden_errors"][:, 1]).max() <= 0.05
    # and it is the same computation as the Python replay through the C-ABI (6 significant digits in the files)
    py = replay.evaluate(dataset, replay.run_icp_test(dataset, backends.GpuBackend()))
    assert np.abs(got["traj"][:, 1:3] - py["traj"][:, 1:3]).max() <= 2e-3
    assert abs(got["traj_scaled"][1, 1] / got["traj"][1, 1] - py["scale"]) <= 1e-4
