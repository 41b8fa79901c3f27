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
    E, R, t, mask, good = oracle.essential_recover(cv2fx["K"], x1, x2, method="8pt")
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


def _ransac_cases(cv2fx, cv2pose):
    out = [(f"{pre}{n}", cv2fx) for pre, n in _cases()]
    out += [(f"c{i}", cv2pose) for i in range(int(cv2pose["n_cases"])) if np.any(cv2pose[f"c{i}_E"])]
    return out


# samples whose degree-10 polynomial has clustered roots: the minimal solution itself is ill-conditioned, so two
# correct solvers agree only to ~1e-5 there (measured: c14 3.7e-6, c11 8.7e-8, c18 7.9e-9); everything else <= 1e-9
_RANSAC_E_TOL = {"c14": 2e-5, "c11": 1e-6, "c18": 1e-7}


def test_ransac_essential_reproduces_cv2(oracle, cv2fx, cv2pose):
    """src/cam.cpp:49 - cv::findEssentialMat(RANSAC) restated (oracle/five_point.cpp: OpenCV's RNG and subset
    sequence, its SVD null-space basis, Nister's solver, solvePoly's root order, Sampson inliers, adaptive iteration
    count) against cv2 4.13.0 itself on 32 problems: the bundled dataset's frame pairs (1 iteration, all inliers),
    synthetic motions with noise and 10 % gross outliers (up to 122 iterations).  The SAME hypothesis must win: E equal
    up to sign to 1e-9, RANSAC inlier mask identical where the fixture has it, then recoverPose's R, t, mask equal."""
    worst = 0.0
    for name, f in _ransac_cases(cv2fx, cv2pose):
        x1, x2, Ecv = f[name + "_x1"], f[name + "_x2"], f[name + "_E"]
        E, rmask, good, iters = oracle.find_essential_ransac(f["K"], x1, x2)
        d = min(np.abs(E - Ecv).max(), np.abs(E + Ecv).max())
        assert d <= _RANSAC_E_TOL.get(name, 1e-9), (name, d)
        worst = max(worst, d if name not in _RANSAC_E_TOL else 0.0)
        assert abs(np.linalg.norm(E) - 1) < 1e-12
        if name + "_ransac_mask" in f:
            assert np.array_equal(rmask != 0, f[name + "_ransac_mask"].ravel() != 0), name
            assert good == int((f[name + "_ransac_mask"].ravel() != 0).sum())
        E2, R, t, mask, g2 = oracle.essential_recover(f["K"], x1, x2, method="ransac")
        tol = 10 * _RANSAC_E_TOL.get(name, 1e-9)
        assert np.abs(R - f[name + "_R"]).max() <= tol and np.abs(t - f[name + "_t"].ravel()).max() <= tol, name
        assert np.array_equal(mask, f[name + "_mask"].ravel()) and g2 == int(f[name + "_good"]), name
    assert worst <= 1e-9


def test_ransac_building_blocks(oracle, cv2fx):
    """the sample sequence depends only on the number of points; every five-point solution is an essential matrix that
    satisfies the five epipolar constraints"""
    sub = oracle.ransac_subsets(115, 50)
    assert sub.shape == (50, 5) and sub.min() >= 0 and sub.max() < 115
    assert all(len(set(r.tolist())) == 5 for r in sub)
    assert sub[0].tolist() == [100, 4, 65, 28, 16]  # cv::RNG((uint64)-1): the sample behind cv2's E on frames 0/1
    assert np.array_equal(sub[:7], oracle.ransac_subsets(115, 7))
    K = cv2fx["K"].astype(np.float64)
    x1, x2 = cv2fx["ds0_x1"].astype(np.float64), cv2fx["ds0_x2"].astype(np.float64)
    q1 = (x1 - K[:2, 2]) / K[[0, 1], [0, 1]]
    q2 = (x2 - K[:2, 2]) / K[[0, 1], [0, 1]]
    Es = oracle.five_point(q1[sub[0]], q2[sub[0]])
    assert 1 <= len(Es) <= 10
    for E in Es:
        s = np.linalg.svd(E, compute_uv=False)
        assert abs(s[0] - s[1]) < 1e-8 and s[2] < 1e-8
        a = np.concatenate([q1[sub[0]], np.ones((5, 1))], 1)
        b = np.concatenate([q2[sub[0]], np.ones((5, 1))], 1)
        assert np.abs(((b @ E) * a).sum(1)).max() < 1e-12


def test_triangulation_kat_vs_world_dat(oracle, dataset, world_gt):
    replay.triangulation_kat(backends.OracleBackend(), dataset, world_gt)


def test_essential_inlier_kat(oracle, dataset):
    """exec/pose_recovery_test.cpp:29-62 as a known-answer test: on the noise-free dataset EVERY match of every
    consecutive frame pair is a RANSAC inlier of the winning hypothesis (1 px threshold), and wherever the robot
    translates (frames 0..50) every match also passes recoverPose's cheirality vote (where it turns on the spot the
    50-baseline distance cap of recoverPose drops far points, as in OpenCV)."""
    for i in range(0, 120, 3):
        a, b = replay.frame(dataset, i), replay.frame(dataset, i + 1)
        m, _ = oracle.match(a["desc"], b["desc"], 0.2, 0.8)
        if len(m) < 8:
            continue
        x1, x2 = a["uv"][m[:, 0]], b["uv"][m[:, 1]]
        E, rmask, rgood, iters = oracle.find_essential_ransac(replay.K_REF, x1, x2)
        assert rgood == len(m) and rmask.all() and iters <= 2, i
        if i <= 50:
            assert oracle.essential_recover(replay.K_REF, x1, x2)[4] == len(m), i


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
    """The whole replay on the oracle alone, the first pose from the restated findEssentialMat(RANSAC) + recoverPose:
    SURVEY 8(c)'s tolerances against output/ - 490 world points with the golden ids, trajectory <= 0.6 % of the 41.4
    extent, heading <= 0.012 rad."""
    res = replay.run_icp_test(dataset, backends.OracleBackend())
    ev, dxy, dth, derr = _replay_metrics(dataset, res)
    assert len(res["world"].xyz) == 490
    assert np.array_equal(ev["world_points"][:, 0], dataset["golden_world_points"][:, 0])
    assert dxy <= 0.006 * 41.4, dxy
    assert dth <= 0.012, dth
    assert derr <= 0.03, derr


def test_replay_full_oracle_linear_estimator(oracle, dataset):
    """Same replay with the batched-sequence option (normalised 8-point on all matches). A different estimator: the
    initial E moves the monocular scale gauge, so the trajectory tolerance is the looser 1 % of extent; the map must
    still be the reference's 490 landmarks."""
    res = replay.run_icp_test(dataset, backends.OracleBackend(essential="8pt"))
    ev, dxy, dth, derr = _replay_metrics(dataset, res)
    assert len(res["world"].xyz) == 490
    assert np.array_equal(ev["world_points"][:, 0], dataset["golden_world_points"][:, 0])
    assert dxy <= 0.01 * 41.4, dxy
    assert dth <= 0.015, dth
    assert derr <= 0.05, derr


def _numpy_match_rows(A, B):
    """match_points' per-row (best, second, idx) restated a second time, independently of oracle/vo_oracle.cpp: numpy
    float32 arithmetic (every operation rounded once), Eigen's squaredNorm order for 10 coefficients
    and the sequential update rule of my_utilities.h:93-99 (strict comparisons, first index wins, NaN / inf ignored)."""
    FLT_MAX = np.finfo(np.float32).max
    with np.errstate(all="ignore"):
        d = A[:, None, :] - B[None, :, :]
        x = d * d
        s = ((x[..., 0] + x[..., 4]) + (x[..., 2] + x[..., 6])) + ((x[..., 1] + x[..., 5]) + (x[..., 3] + x[..., 7]))
        s = (s + x[..., 8]) + x[..., 9]
    assert s.dtype == np.float32
    usable = s < FLT_MAX                      # NaN and values >= FLT_MAX never pass `d < best`
    key = np.where(usable, s, np.float32(np.inf))
    idx = np.argmin(key, axis=1)              # first minimum
    best = key[np.arange(len(A)), idx]
    none = ~usable.any(axis=1)
    masked = key.copy()
    masked[np.arange(len(A)), idx] = np.inf
    second = masked.min(axis=1)
    best = np.where(none, FLT_MAX, best).astype(np.float32)
    second = np.where(np.isinf(second), FLT_MAX, second).astype(np.float32)
    return best, second, np.where(none, -1, idx).astype(np.int32)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_oracle_matcher_against_an_independent_numpy_restatement(oracle, seed):
    """the oracle is what the CUDA matcher is compared with bit for bit; here it is itself compared bit for bit with a
    second restatement on adversarial sets (ties, zero distances, huge / tiny scales, offsets, NaN, inf)"""
    import synth
    rng = np.random.default_rng(seed)
    for kind in ("uniform", "clustered", "lattice", "lowrank", "cauchy"):
        A, B = synth.stress_descriptors(rng, kind, int(rng.integers(1, 300)), int(rng.integers(1, 700)))
        if rng.random() < 0.5:
            A[rng.integers(0, len(A)), rng.integers(0, 10)] = np.nan
            B[rng.integers(0, len(B)), rng.integers(0, 10)] = np.inf
        _, _, best, second, idx = oracle.match(A, B, want_rows=True)
        nb, ns, ni = _numpy_match_rows(A, B)
        assert np.array_equal(idx, ni), kind
        assert np.array_equal(best.view(np.uint32), nb.view(np.uint32)), kind
        assert np.array_equal(second.view(np.uint32), ns.view(np.uint32)), kind


def _numpy_picp_status(K, rows, cols, pose, world, image, pairs, thr):
    """Camera::projectPoint + the chi test of PICPSolver::linearize (camera.h:24-36, picp_solver.cpp:65-80) restated a
    second time in numpy float32: length-3 products reduce as x0 + (x1 + x2), the reciprocal is formed in double and
    rounded to float, every comparison keeps the reference's NaN behaviour (a NaN never leaves the image)."""
    f = np.float32
    K = K.astype(f); T = pose.astype(f)
    p = world[pairs[:, 1]].astype(f); z = image[pairs[:, 0]].astype(f)
    with np.errstate(all="ignore"):
        dot3 = lambda a0, b0, a1, b1, a2, b2: (a0 * b0 + (a1 * b1 + a2 * b2)).astype(f)
        c = [(T[i, 3] + dot3(T[i, 0], p[:, 0], T[i, 1], p[:, 1], T[i, 2], p[:, 2])).astype(f) for i in range(3)]
        q = [dot3(K[i, 0], c[0], K[i, 1], c[1], K[i, 2], c[2]) for i in range(3)]
        iz = (1.0 / q[2].astype(np.float64)).astype(f)
        u, v = (q[0] * iz).astype(f), (q[1] * iz).astype(f)
        outside = (c[2] <= 0) | (u < 0) | (u > f(cols - 1)) | (v < 0) | (v > f(rows - 1))
        e0, e1 = (u - z[:, 0]).astype(f), (v - z[:, 1]).astype(f)
        chi = (e0 * e0 + e1 * e1).astype(f)
        st = np.where(outside, 0, np.where(chi > f(thr), 2, 1)).astype(np.uint8)  # VO_PICP_SKIPPED / OUTLIER / INLIER
    return st


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_picp_status_against_an_independent_numpy_restatement(oracle, seed):
    import synth
    rng = np.random.default_rng(seed)
    for _ in range(8):
        K, rows, cols, pose, world, image, pairs, thr, keep, general, scale = synth.picp_stress_case(rng)
        ref = oracle.linearize(K, rows, cols, pose, world, image, pairs, thr, keep, accum="f64")
        st = _numpy_picp_status(K, rows, cols, pose, world, image, pairs, thr)
        assert np.array_equal(ref["status"], st), (len(pairs), scale, general, thr)


def test_oracle_isometry_algebra_against_numpy_float32(oracle):
    """Isometry3f inverse and product (exec/icp_test.cpp:79,114; Eigen: coefficient products reduced as x0 + (x1 + x2)),
    restated in numpy float32 and compared bit for bit"""
    f = np.float32
    rng = np.random.default_rng(0)
    dot3 = lambda a0, b0, a1, b1, a2, b2: f(f(a0 * b0) + f(f(a1 * b1) + f(a2 * b2)))
    for _ in range(200):
        scale = f(rng.choice([1e-3, 1.0, 50.0, 1e6]))
        A = np.concatenate([rng.normal(0, 1, (3, 3)), rng.normal(0, 1, (3, 1)) * scale], 1).astype(f)
        B = np.concatenate([rng.normal(0, 1, (3, 3)), rng.normal(0, 1, (3, 1)) * scale], 1).astype(f)
        inv = np.zeros((3, 4), f)
        for i in range(3):
            inv[i, :3] = A[:3, i]
            inv[i, 3] = dot3(-A[0, i], A[0, 3], -A[1, i], A[1, 3], -A[2, i], A[2, 3])
        assert np.array_equal(oracle.pose_inverse(A).view(np.uint32), inv.view(np.uint32))
        mul = np.zeros((3, 4), f)
        for i in range(3):
            for j in range(4):
                mul[i, j] = dot3(A[i, 0], B[0, j], A[i, 1], B[1, j], A[i, 2], B[2, j])
            mul[i, 3] = f(mul[i, 3] + A[i, 3])
        assert np.array_equal(oracle.pose_mul(A, B).view(np.uint32), mul.view(np.uint32))


def _tri_stats(X, R, X4):
    """relative error per point, over the points whose homogeneous coordinate is not vanishing (|w| >= 1e-4 max|X4|:
    dividing by a w that is pure rounding noise is not a comparison of implementations)"""
    with np.errstate(all="ignore"):
        rel = np.abs(X - R).max(1) / np.maximum(np.abs(R).max(1), 1e-30)
        ok = np.isfinite(rel) & (np.abs(X4[:, 3]) >= 1e-4 * np.abs(X4).max(1))
    rel = rel[ok]
    return float(np.median(rel)), float(np.quantile(rel, 0.9))


def test_triangulate_vs_cv2_on_adversarial_two_view_problems(oracle, cv2tri):
    """both cameras away from the origin, scene scales 0.01..100, baselines 1e-4..2 of the scene scale, 0..2 px noise,
    points almost at infinity: against cv2 4.13's triangulatePoints the restated DLT agrees to float32 output
    precision on the bulk of every case (ill-conditioned points - zero parallax, w -> 0 - amplify the last-bit
    differences of the two SVD implementations and are left to the 90 % quantile)"""
    K = cv2tri["K"]
    for c in range(int(cv2tri["n_cases"])):
        X = oracle.triangulate(K, cv2tri[f"c{c}_T1"], cv2tri[f"c{c}_T2"], cv2tri[f"c{c}_x1"], cv2tri[f"c{c}_x2"])
        med, q90 = _tri_stats(X, cv2tri[f"c{c}_X3"], cv2tri[f"c{c}_X4"])
        assert med <= 1e-4 and q90 <= 2e-3, (c, cv2tri[f"c{c}_cfg"], med, q90)


def test_recover_pose_matches_cv2_on_more_motions(oracle, cv2pose):
    """sideways / forward / backward baselines, rotations up to ~0.5 rad, 0..1 px noise, 10 % gross outliers,
    20..400 points: given cv2's E, the restated decomposeEssentialMat + cheirality vote returns cv2's R, t, mask
    and inlier count (src/cam.cpp:61)"""
    K = cv2pose["K"]
    for c in range(int(cv2pose["n_cases"])):
        E = cv2pose[f"c{c}_E"]
        if not E.any():
            continue
        R, t, mask, good = oracle.recover_pose(E, K, cv2pose[f"c{c}_x1"], cv2pose[f"c{c}_x2"])
        assert np.abs(R - cv2pose[f"c{c}_R"]).max() < 1e-12, c
        assert np.abs(t - cv2pose[f"c{c}_t"]).max() < 1e-12, c
        assert np.array_equal(mask > 0, cv2pose[f"c{c}_mask"] > 0), c
        assert int(good) == int(cv2pose[f"c{c}_good"]), c
