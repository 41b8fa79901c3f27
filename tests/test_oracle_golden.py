"""Pins the CPU oracle: against cv2 4.13 black-box fixtures (tests/golden/cv2_fixtures.npz,
made by oracle/gen_golden.py) and against the reference's own end-to-end goldens
(output/*.txt, carried in tests/golden/dataset.npz).  CPU only."""
import numpy as np
import pytest

import backends
import replay

DS_PAIRS, SYN = 7, 5
I34 = np.eye(4, dtype=np.float32)[:3].copy()


def _cases():
    return [("ds", n) for n in range(DS_PAIRS)] + [("syn", n) for n in range(SYN)]


@pytest.mark.parametrize("pre,n", _cases())
def test_recover_pose_matches_cv2(oracle, cv2fx, pre, n):
    """OpenCV recoverPose restated (src/cam.cpp:61): given cv2's E, same R, t, mask."""
    x1, x2 = cv2fx[f"{pre}{n}_x1"], cv2fx[f"{pre}{n}_x2"]
    R, t, mask, good = oracle.recover_pose(cv2fx[f"{pre}{n}_E"], cv2fx["K"], x1, x2)
    assert np.abs(R - cv2fx[f"{pre}{n}_R"]).max() < 1e-12
    assert np.abs(t - cv2fx[f"{pre}{n}_t"]).max() < 1e-12
    assert np.array_equal(mask > 0, cv2fx[f"{pre}{n}_mask"] > 0)
    assert good == int(cv2fx[f"{pre}{n}_good"])


@pytest.mark.parametrize("pre,n", _cases())
def test_triangulate_matches_cv2(oracle, cv2fx, pre, n):
    """cv::triangulatePoints + convertPointsFromHomogeneous (src/cam.cpp:108-118). Tolerance:
    1e-4 relative to the cloud extent (float32 output of a float64 DLT)."""
    x1, x2 = cv2fx[f"{pre}{n}_x1"], cv2fx[f"{pre}{n}_x2"]
    T2 = oracle.pose_inverse(cv2fx[f"{pre}{n}_T2inv"][:3])
    X = oracle.triangulate(cv2fx["K"], I34, T2, x1, x2)
    ref = cv2fx[f"{pre}{n}_X3"]
    assert np.abs(X - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("pre,n", _cases())
def test_essential_vs_cv2_blackbox(oracle, cv2fx, pre, n):
    """8-point on all matches vs cv2's un-refitted minimal-sample 5-point E: a black-box
    comparison (different estimators). Noise-free: R <= 1e-4, t <= 2e-3; noisy synthetic:
    both must be closer than 1e-2 and the 8-point must not be further from GT than cv2."""
    x1, x2 = cv2fx[f"{pre}{n}_x1"], cv2fx[f"{pre}{n}_x2"]
    E, R, t, mask, good = oracle.essential_recover(cv2fx["K"], x1, x2)
    assert good == len(x1)
    noise = 0.0 if pre == "ds" else float(cv2fx["syn_cfg"][n][2])
    dR = np.abs(R - cv2fx[f"{pre}{n}_R"]).max()
    dt = np.abs(t - cv2fx[f"{pre}{n}_t"]).max()
    if noise == 0.0:
        assert dR < 1e-4 and dt < 2e-3
    else:
        assert dR < 1e-2 and dt < 1e-2
        Rgt, tgt = cv2fx[f"{pre}{n}_Rgt"], cv2fx[f"{pre}{n}_tgt"]
        assert np.abs(R - Rgt).max() <= np.abs(cv2fx[f"{pre}{n}_R"] - Rgt).max() + 1e-6
        assert np.abs(t - tgt).max() <= np.abs(cv2fx[f"{pre}{n}_t"] - tgt).max() + 1e-6
    # essential-manifold properties
    s = np.linalg.svd(E, compute_uv=False)
    assert abs(s[0] - 1) < 1e-9 and abs(s[1] - 1) < 1e-9 and s[2] < 1e-9
    assert abs(np.linalg.det(R) - 1) < 1e-9 and abs(np.linalg.norm(t) - 1) < 1e-9


def test_matching_kat_ids(oracle, dataset):
    """exec/match_points_test.cpp:20-39 as a KAT: on data/ every accepted pair has equal id_real,
    and every id present in both frames is found (descriptors are exact copies)."""
    for i in range(0, 120, 7):
        a, b = replay.frame(dataset, i), replay.frame(dataset, i + 1)
        pairs, (possible, correct) = oracle.match(a["desc"], b["desc"], 0.2, 0.8, a["id_real"], b["id_real"])
        assert correct == len(pairs)
        common = len(set(a["id_real"].tolist()) & set(b["id_real"].tolist()))
        assert possible == common == len(pairs)
        assert np.array_equal(a["id_real"][pairs[:, 0]], b["id_real"][pairs[:, 1]])
        assert np.all(np.diff(pairs[:, 0]) > 0)


def _replay_metrics(dataset, res):
    ev = replay.evaluate(dataset, res)
    g = dataset
    dxy = np.linalg.norm(ev["traj"][:, 1:3] - g["golden_traj"][:, 1:3], axis=1).max()
    dth = np.abs(ev["traj"][:, 3] - g["golden_traj"][:, 3]).max()
    derr = np.abs(ev["errors"][:, 1] - g["golden_errors"][:, 1]).max()
    return ev, dxy, dth, derr


def test_replay_anchored_on_cv2_pose_reproduces_output(oracle, dataset, cv2fx):
    """The reference's only pinned results (output/*.txt from exec/icp_test.cpp). With cv2-4.13's
    (R,t) for frames 0/1 as the initial pose, everything downstream (matching, triangulation,
    PICP, anti-join) is the oracle.  Tolerances from SURVEY 8(c)/BASELINE.md: 490 world points
    exactly, trajectory <= 0.6 % of the 41.4 extent, heading <= 0.012 rad."""
    be = backends.OracleBackend(cv2_first_pose=(cv2fx["ds0_R"], cv2fx["ds0_t"], cv2fx["ds0_mask"]))
    res = replay.run_icp_test(dataset, be)
    ev, dxy, dth, derr = _replay_metrics(dataset, res)
    assert res["n_init_matches"] == 115
    assert len(res["world"].xyz) == 490 and len(ev["world_points"]) == 490
    assert np.array_equal(ev["world_points"][:, 0], dataset["golden_world_points"][:, 0])
    assert dxy <= 0.006 * 41.4, dxy
    assert dth <= 0.012, dth
    assert derr <= 0.03, derr
    golden_scale = dataset["golden_traj_scaled"][1, 1] / dataset["golden_traj"][1, 1]
    assert abs(ev["scale"] - golden_scale) < 5e-4


def test_replay_full_oracle(oracle, dataset):
    """Same replay with the oracle's own essential estimator (8-point on all matches). The
    different initial E moves the monocular scale gauge, so the trajectory tolerance is the
    looser 1 % of extent; the map must still be the reference's 490 landmarks."""
    res = replay.run_icp_test(dataset, backends.OracleBackend())
    ev, dxy, dth, derr = _replay_metrics(dataset, res)
    assert len(res["world"].xyz) == 490
    assert np.array_equal(ev["world_points"][:, 0], dataset["golden_world_points"][:, 0])
    assert dxy <= 0.01 * 41.4, dxy
    assert dth <= 0.015, dth
    assert derr <= 0.05, derr
