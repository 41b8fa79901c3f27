"""GPU parity tests for the PICP path (through the C-ABI) against the CPU oracle.

Bars (BASELINE.json north_star): inlier masks bit-exact; H/b <= 1e-4 norm-relative against the
float64-accumulated oracle; pose <= 1e-5 absolute after the same number of rounds."""
import numpy as np
import pytest

import synth
from backends import product

pytestmark = pytest.mark.gpu

H_TOL = 1e-4
POSE_TOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    vo = product()
    c = vo.Context(0)
    yield c
    c.close()


def _solver(ctx, fr, pose=None, mode=0):
    s = ctx.picp()
    s.set_mode(mode)
    s.set_camera(fr["K"], fr["rows"], fr["cols"], fr["pose0"] if pose is None else pose)
    s.set_points(fr["world"], fr["image"])
    s.set_correspondences(fr["pairs"])
    return s


def _check_lin(lin, ref):
    assert np.array_equal(lin["status"], ref["status"])
    assert lin["n_inliers"] == ref["n_inliers"]
    assert lin["n_outliers"] == int((ref["status"] == 2).sum())
    hs = max(np.abs(ref["H"]).max(), 1e-30)
    bs = max(np.abs(ref["b"]).max(), 1e-30)
    assert np.abs(lin["H"] - ref["H"]).max() <= H_TOL * hs
    assert np.abs(lin["b"] - ref["b"]).max() <= H_TOL * bs
    assert abs(lin["chi_in"] - ref["chi_in"]) <= H_TOL * max(ref["chi_in"], 1.0)
    assert abs(lin["chi_out"] - ref["chi_out"]) <= H_TOL * max(ref["chi_out"], 1.0)
    assert np.array_equal(lin["H"], lin["H"].T)


@pytest.mark.parametrize("n,permute", [(1, False), (3, False), (4, False), (5, False), (257, False), (1000, True),
                                       (65536, False), (200003, True)])
@pytest.mark.parametrize("thr,keep", [(3000.0, False), (100.0, True), (1000.0, False)])
def test_linearize_matches_oracle(ctx, oracle, n, permute, thr, keep):
    fr = synth.picp_frame(n=n, seed=100 + n, permute=permute)
    s = _solver(ctx, fr)
    lin = s.linearize(thr, keep, want_status=True, n_pairs=n)
    ref = oracle.linearize(fr["K"], fr["rows"], fr["cols"], fr["pose0"], fr["world"], fr["image"], fr["pairs"], thr,
                           keep, accum="f64")
    _check_lin(lin, ref)
    s.close()


def test_linearize_general_camera_matrix(ctx, oracle):
    """non-pinhole K (skew, K[2][2] != 1) takes the general arithmetic path"""
    fr = synth.picp_frame(n=5000, seed=5)
    K = np.array([[181.5, 0.7, 318.2], [0.01, 179.3, 241.1], [1e-4, -2e-4, 1.01]], np.float32)
    fr["K"] = K
    s = _solver(ctx, fr)
    lin = s.linearize(3000.0, False, want_status=True, n_pairs=5000)
    ref = oracle.linearize(K, fr["rows"], fr["cols"], fr["pose0"], fr["world"], fr["image"], fr["pairs"], 3000.0, False,
                           accum="f64")
    _check_lin(lin, ref)
    s.close()


def test_linearize_edge_values(ctx, oracle):
    """z<=0, exactly-on-border pixels, chi exactly at the threshold, NaN/inf coordinates."""
    K = synth.K_REF
    I = np.eye(4, dtype=np.float32)[:3]
    world = np.array([[0, 0, 1], [0, 0, 0], [0, 0, -1], [-320 / 180, 0, 1], [319 / 180, 0, 1], [319.5 / 180, 0, 1],
                      [0, -240 / 180, 1], [0, 239 / 180, 1], [0, 240 / 180, 1], [np.nan, 0, 1], [0, np.inf, 1],
                      [0, 0, np.inf], [0, 0, np.nan], [1e-30, 1e-30, 1e-38], [0.5, 0.25, 2.0]], np.float32)
    n = len(world)
    image = np.zeros((n, 2), np.float32)
    image[:] = [320, 240]
    image[0] = [320 + 30, 240 + 40]  # chi = 2500 exactly
    image[14] = [320 + 45 + 30, 240 + 22.5 + 40]
    pairs = np.stack([np.arange(n), np.arange(n)], 1).astype(np.int32)
    for thr in (2500.0, 2499.9998, 3000.0):
        for keep in (False, True):
            s = ctx.picp()
            s.set_camera(K, 480, 640, I)
            s.set_points(world, image)
            s.set_correspondences(pairs)
            lin = s.linearize(thr, keep, want_status=True, n_pairs=n)
            ref = oracle.linearize(K, 480, 640, I, world, image, pairs, thr, keep, accum="f64")
            assert np.array_equal(lin["status"], ref["status"]), (thr, keep, lin["status"], ref["status"])
            assert lin["n_inliers"] == ref["n_inliers"]
            s.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_linearize_adversarial_sweep(ctx, oracle, seed):
    """random cameras (pinhole / general K), any rotation, scene scales 1e-3..1e6, points behind and on the camera
    plane, border-pixel measurements, NaN / inf / overflowing points, thresholds 1e-6..1e12 (exp/picp_stress.py):
    the per-correspondence status is bit-exact, H and b within 1e-4 whenever the reference's are finite"""
    rng = np.random.default_rng(seed)
    for case in range(12):
        K, rows, cols, pose, world, image, pairs, thr, keep, general, scale = synth.picp_stress_case(rng)
        s = ctx.picp()
        s.set_camera(K, rows, cols, pose)
        s.set_points(world, image)
        s.set_correspondences(pairs)
        lin = s.linearize(thr, keep, want_status=True, n_pairs=len(pairs))
        ref = oracle.linearize(K, rows, cols, pose, world, image, pairs, thr, keep, accum="f64")
        s.close()
        tag = f"seed {seed} case {case} n {len(pairs)} scale {scale:g} general {general} thr {thr:g} keep {keep}"
        assert np.array_equal(lin["status"], ref["status"]), tag
        assert lin["n_inliers"] == ref["n_inliers"], tag
        with np.errstate(all="ignore"):
            hs, bs = np.abs(ref["H"]).max(), np.abs(ref["b"]).max()
        if np.isfinite(hs) and np.isfinite(bs):
            assert np.abs(lin["H"] - ref["H"]).max() <= 1e-4 * hs, tag
            assert np.abs(lin["b"] - ref["b"]).max() <= 1e-4 * max(bs, 1e-30), tag
        else:  # an inlier with non-finite terms poisons the system in the reference; it must do so here too
            assert not (np.isfinite(lin["H"]).all() and np.isfinite(lin["b"]).all()), tag


@pytest.mark.parametrize("keep", [False, True])
def test_dropped_points_with_non_finite_terms_leave_the_sums_finite(ctx, oracle, keep):
    """A point that does not contribute is removed by zeroing its 1/z alone - exact only while its other terms are
    finite.  Points that are dropped AND carry non-finite terms (coordinates around 3e38 whose K c overflows: skipped;
    an inf measurement: - without keep_outliers - a rejected outlier) must take the zero-everything accumulation:
    the statuses are the oracle's and H, b stay finite and within 1e-4 of the oracle's, in linearize and in all three
    round kernels; sprinkled so that every quad position and both pairs of a quad are hit."""
    fr = synth.picp_frame(n=6000, seed=77)
    world, image = fr["world"].copy(), fr["image"].copy()
    # finite coordinates whose q = K c overflows (an inf coordinate would turn into a NaN depth through 0 * inf and
    # poison the reference's system as an inlier: that case is in the adversarial sweep above)
    bad_world = np.array([[3e38, 0, 5], [-3e38, 1, 5], [0, 3e38, 5], [2e38, 2e38, 5], [1, 1, -3e38], [3e38, 3e38, -3e38],
                          [1e37, -1e37, 1e-3]], np.float32)
    for j, k in enumerate(range(3, 6000, 97)):
        world[fr["pairs"][k, 1]] = bad_world[(j + k) % len(bad_world)]
    for k in range(50, 6000, 211):  # a finite projection against an infinite measurement: chi = inf > thr
        image[fr["pairs"][k, 0], k % 2] = np.inf if k % 3 else -np.inf
    thr = 3000.0
    ref = oracle.linearize(fr["K"], fr["rows"], fr["cols"], fr["pose0"], world, image, fr["pairs"], thr, keep, accum="f64")
    with np.errstate(all="ignore"):
        ref_finite = np.isfinite(ref["H"]).all() and np.isfinite(ref["b"]).all()
    assert ref_finite == (not keep)  # a KEPT outlier with chi = inf poisons the reference's system (lambda = 0, e = inf)
    s = ctx.picp()
    s.set_camera(fr["K"], fr["rows"], fr["cols"], fr["pose0"])
    s.set_points(world, image)
    s.set_correspondences(fr["pairs"])
    lin = s.linearize(thr, keep, want_status=True, n_pairs=len(fr["pairs"]))
    assert np.array_equal(lin["status"], ref["status"])
    assert (ref["status"] == 0).sum() >= 60 and (ref["status"] == 2).sum() >= 20
    assert lin["n_inliers"] == ref["n_inliers"] and lin["n_outliers"] == int((ref["status"] == 2).sum())
    if ref_finite:
        assert np.isfinite(lin["H"]).all() and np.isfinite(lin["b"]).all()
        assert np.abs(lin["H"] - ref["H"]).max() <= H_TOL * np.abs(ref["H"]).max()
        assert np.abs(lin["b"] - ref["b"]).max() <= H_TOL * np.abs(ref["b"]).max()
        assert abs(lin["chi_in"] - ref["chi_in"]) <= H_TOL * max(ref["chi_in"], 1.0)
        assert lin["chi_out"] == np.inf
    else:
        assert not (np.isfinite(lin["H"]).all() and np.isfinite(lin["b"]).all())
    s.close()
    if not ref_finite:
        return
    poses = []
    for mode in (1, 2, 3):  # per-round launches, shared-memory resident, persistent streaming
        s = ctx.picp()
        s.set_mode(mode)
        s.set_camera(fr["K"], fr["rows"], fr["cols"], fr["pose0"])
        s.set_points(world, image)
        s.set_correspondences(fr["pairs"])
        s.enqueue_rounds(thr, 1.0, keep, 4)
        st = s.fetch_stats(4)
        assert st[0].num_inliers == ref["n_inliers"], mode
        poses.append(s.get_pose())
        assert np.isfinite(poses[-1]).all(), mode
        s.close()
    pose = np.array(fr["pose0"], np.float32)
    for _ in range(4):
        pose, _, _, _ = oracle.one_round(fr["K"], fr["rows"], fr["cols"], pose, world, image, fr["pairs"], thr, 1.0, keep)
    for q in poses:
        assert np.abs(q - pose).max() <= POSE_TOL


def test_mask_on_threshold_knife_edge(ctx, oracle):
    """Measurements placed so that chi lands within a few ulps of the kernel threshold for EVERY
    correspondence: the inlier mask then depends on the last bit of the projection (the kernel's
    hand-rolled reciprocal and pinhole shortcut must be bit-identical to the reference arithmetic)."""
    fr = synth.picp_frame(n=150000, seed=77, outlier_frac=0.0, invalid_frac=0.0)
    uv, _ = oracle.project_points(fr["K"], 480, 640, fr["pose0"], fr["world"], keep_indices=True)
    rng = np.random.default_rng(3)
    ang = rng.uniform(0, 2 * np.pi, len(uv))
    r = 50.0 * (1 + rng.integers(-3, 4, len(uv)) * 2.0 ** -23)
    fr["image"] = (uv + np.stack([r * np.cos(ang), r * np.sin(ang)], 1)).astype(np.float32)
    s = _solver(ctx, fr)
    lin = s.linearize(2500.0, False, want_status=True, n_pairs=len(uv))
    ref = oracle.linearize(fr["K"], 480, 640, fr["pose0"], fr["world"], fr["image"], fr["pairs"], 2500.0, False,
                           accum="f64")
    assert np.array_equal(lin["status"], ref["status"])
    frac_in = (ref["status"] == 1).mean()
    assert 0.2 < frac_in < 0.8  # the threshold really cuts through the set
    s.close()


def test_empty_and_invalid_correspondences(ctx):
    vo = product()
    fr = synth.picp_frame(n=16, seed=1)
    s = ctx.picp()
    s.set_camera(fr["K"], 480, 640, fr["pose0"])
    s.set_points(fr["world"], fr["image"])
    s.set_correspondences(np.zeros((0, 2), np.int32))
    lin = s.linearize(3000.0, False)
    assert lin["n_inliers"] == 0 and not lin["H"].any() and not lin["b"].any()
    st = s.one_round(3000.0, 1.0, False)  # H = I, b = 0 -> dx = 0 -> pose unchanged
    assert st.num_inliers == 0
    assert np.array_equal(s.get_pose(), fr["pose0"])
    bad = np.array([[0, 0], [3, 99]], np.int32)
    with pytest.raises(vo.VoError):
        s.set_correspondences(bad)
    with pytest.raises(vo.VoError):  # state invalidated by the failed call
        s.one_round(3000.0, 1.0, False)
    s.close()
    s2 = ctx.picp()
    with pytest.raises(vo.VoError):
        s2.one_round(3000.0, 1.0, False)  # nothing initialised
    s2.close()


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["stream", "resident", "stream_persistent"])
@pytest.mark.parametrize("damping", [0.0, 1e-3, 250.0])
def test_damping_values_track_oracle(ctx, oracle, damping, mode):
    """H += I * damping (picp_solver.cpp:96): damping 0 takes the pivoted LDLT restatement (one lane, Eigen's
    algorithm) instead of the unpivoted factorisation every other test runs; small and large damping stay on the
    unpivoted one. Every kernel's solve against the oracle's, 6 rounds."""
    n = 20000
    fr = synth.picp_frame(n=n, seed=99)
    s = _solver(ctx, fr, mode=mode)
    pose = fr["pose0"].copy()
    s.enqueue_rounds(3000.0, damping, False, 6)
    batch = s.fetch_stats(6)
    for r in range(6):
        pose, ci, co, ni = oracle.one_round(fr["K"], fr["rows"], fr["cols"], pose, fr["world"], fr["image"], fr["pairs"],
                                            3000.0, damping, False)
        assert abs(batch[r].num_inliers - ni) <= (0 if r == 0 else 2), (damping, r)
        assert abs(batch[r].chi_inliers - ci) <= 2e-4 * max(ci, 1.0), (damping, r)
    assert np.abs(s.get_pose() - pose).max() <= 2 * POSE_TOL, damping
    s.close()


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["stream", "resident", "stream_persistent"])
@pytest.mark.parametrize("n,permute,thr,keep,rounds", [(50000, False, 3000.0, False, 10), (50000, True, 100.0, True, 10),
                                                       (120, False, 3000.0, False, 8), (300000, False, 1000.0, False, 5)])
def test_rounds_track_oracle(ctx, oracle, n, permute, thr, keep, rounds, mode):
    """oneRound x rounds: same inlier counts every round, pose within 1e-5 of the float32-sequential
    oracle, both through the synchronous call and through the enqueue/fetch path - on the streaming kernel
    (one launch per round) and on the resident kernel (all rounds in one persistent launch)."""
    fr = synth.picp_frame(n=n, seed=n + 1, permute=permute)
    s = _solver(ctx, fr, mode=mode)
    s2 = _solver(ctx, fr, mode=mode)
    pose = fr["pose0"].copy()
    s2.enqueue_rounds(thr, 1.0, keep, rounds)
    batch = s2.fetch_stats(rounds)
    for r in range(rounds):
        st = s.one_round(thr, 1.0, keep)
        pose, ci, co, ni = oracle.one_round(fr["K"], fr["rows"], fr["cols"], pose, fr["world"], fr["image"], fr["pairs"],
                                            thr, 1.0, keep)
        # round 0 starts from the same pose: counts are exact. Later rounds run from poses that agree
        # only to ~1e-7, so a correspondence whose chi sits on the threshold may flip.
        slack = 0 if r == 0 else max(2, int(1e-5 * n))
        assert abs(st.num_inliers - ni) <= slack, r
        assert batch[r].num_inliers == st.num_inliers
        assert batch[r].chi_inliers == st.chi_inliers  # deterministic: two runs are bit-identical
        assert abs(st.chi_inliers - ci) <= 2e-4 * max(ci, 1.0)
        assert np.abs(s.get_pose() - pose).max() <= POSE_TOL, r
    assert np.array_equal(s.get_pose(), s2.get_pose())
    s.close()
    s2.close()


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["stream", "resident", "stream_persistent"])
def test_device_side_convergence_loop(ctx, oracle, mode):
    """vo_picp_solve == the driver loop of exec/icp_test.cpp:88-107 run with synchronous rounds."""
    fr = synth.picp_frame(n=20000, seed=11)
    s = _solver(ctx, fr, mode=mode)
    done, last = s.solve(3000.0, 1.0, False, max_rounds=50, rel_tol=1e-3)
    s2 = _solver(ctx, fr, mode=mode)
    prev = np.float32(np.finfo(np.float32).max)
    it = 0
    for it in range(1, 51):
        st = s2.one_round(3000.0, 1.0, False)
        cur = np.float32(st.chi_inliers)
        rel = np.float32(abs(prev - cur)) / prev
        if rel < np.float32(1e-3):
            break
        prev = cur
    assert done == it
    assert last.chi_inliers == st.chi_inliers and last.num_inliers == st.num_inliers
    assert np.array_equal(s.get_pose(), s2.get_pose())
    s.close()
    s2.close()


def test_device_resident_points(ctx, oracle):
    """set_points_dev / set_correspondences_dev borrow HBM buffers (torch is only the allocator)."""
    import torch
    fr = synth.picp_frame(n=30000, seed=3, permute=True)
    dw = torch.from_numpy(fr["world"]).cuda()
    di = torch.from_numpy(fr["image"]).cuda()
    dp = torch.from_numpy(fr["pairs"]).cuda()
    torch.cuda.synchronize()
    s = ctx.picp()
    s.set_camera(fr["K"], 480, 640, fr["pose0"])
    s.set_points_dev(dw.data_ptr(), len(fr["world"]), di.data_ptr(), len(fr["image"]))
    s.set_correspondences_dev(dp.data_ptr(), len(fr["pairs"]))
    lin = s.linearize(3000.0, False, want_status=True, n_pairs=len(fr["pairs"]))
    ref = oracle.linearize(fr["K"], 480, 640, fr["pose0"], fr["world"], fr["image"], fr["pairs"], 3000.0, False,
                           accum="f64")
    _check_lin(lin, ref)
    s.close()


def _sizes_around_resident_geometry(cap):
    # 1 CTA / 2 CTAs boundary (1536 quads-threads x 4), ragged tails, one wave, the capacity itself
    return [1, 3, 5, 1535, 1536, 1537, 6143, 6145, 50001, 148 * 1536 + 7, 1 << 20, cap - 3, cap]


def test_resident_matches_streaming(ctx, oracle):
    """The persistent shared-memory-resident kernel against the one-launch-per-round kernel on the same frames:
    identical inlier / outlier counts in round 0 (same arithmetic on the exact part), every later round within the
    knife-edge slack, chi within 1e-5, final pose within 1e-6; and round 0 against the oracle's counts."""
    probe = ctx.picp()
    cap = probe.resident_capacity
    probe.close()
    assert cap >= 1310720  # BASELINE config 3's 10M frame over 8 GPUs must fit
    for n in _sizes_around_resident_geometry(cap):
        permute = n % 2 == 1
        thr, keep = ((3000.0, False), (100.0, True))[(n // 3) % 2]
        fr = synth.picp_frame(n=n, seed=7 + n, permute=permute)
        res = {}
        for mode in (1, 2):
            s = _solver(ctx, fr, mode=mode)
            s.enqueue_rounds(thr, 1.0, keep, 6)
            res[mode] = (s.fetch_stats(6), s.get_pose())
            s.close()
        (st1, p1), (st2, p2) = res[1], res[2]
        ref = oracle.linearize(fr["K"], 480, 640, fr["pose0"], fr["world"], fr["image"], fr["pairs"], thr, keep, accum="f64")
        assert st2[0].num_inliers == ref["n_inliers"] == st1[0].num_inliers, n
        assert st2[0].num_outliers == st1[0].num_outliers, n
        for r in range(6):
            slack = 0 if r == 0 else max(2, int(1e-5 * n))
            assert abs(st1[r].num_inliers - st2[r].num_inliers) <= slack, (n, r)
            assert abs(st1[r].chi_inliers - st2[r].chi_inliers) <= 1e-5 * max(st1[r].chi_inliers, 1.0), (n, r)
        assert np.abs(p1 - p2).max() <= 1e-6, n


@pytest.mark.parametrize("n", [1, 1407, 1409, 148 * 1408 + 5, 3_000_001])
def test_persistent_streaming_matches_per_round_launches(ctx, n):
    """the persistent streaming kernel (all rounds in one cooperative launch, planes streamed through the TMA ring
    every round) against one launch per round on the same packed planes: the exact part is the same code, the sums
    are reduced in a different (fixed) order - counts identical in round 0, within the knife-edge slack later, chi
    1e-5, pose 1e-6; two runs are bit-identical; the in-kernel convergence test stops at the same round."""
    fr = synth.picp_frame(n=n, seed=3 + n, permute=(n % 2 == 1))
    res = {}
    for mode in (1, 3, 3):
        s = _solver(ctx, fr, mode=mode)
        s.enqueue_rounds(3000.0, 1.0, False, 7)
        res.setdefault(mode, []).append((s.fetch_stats(7), s.get_pose()))
        s.close()
    (st1, p1), (st3, p3), (st3b, p3b) = res[1][0], res[3][0], res[3][1]
    assert np.array_equal(p3, p3b) and [x.chi_inliers for x in st3] == [x.chi_inliers for x in st3b]
    assert st1[0].num_inliers == st3[0].num_inliers and st1[0].num_outliers == st3[0].num_outliers
    for r in range(7):
        assert abs(st1[r].num_inliers - st3[r].num_inliers) <= (0 if r == 0 else max(2, int(1e-5 * n))), r
        assert abs(st1[r].chi_inliers - st3[r].chi_inliers) <= 1e-5 * max(st1[r].chi_inliers, 1.0), r
    assert np.abs(p1 - p3).max() <= 1e-6
    if n >= 1000:
        done = {}
        for mode in (1, 3):
            s = _solver(ctx, fr, mode=mode)
            done[mode] = s.solve(3000.0, 1.0, False, max_rounds=40, rel_tol=1e-3)[0]
            s.close()
        assert done[1] == done[3]


def test_resident_capacity_and_fallback(ctx):
    """a set one correspondence above the resident capacity: AUTO streams it, RESIDENT refuses it"""
    vo = product()
    probe = ctx.picp()
    cap = probe.resident_capacity
    probe.close()
    fr = synth.picp_frame(n=cap + 1, seed=5)
    s = _solver(ctx, fr, mode=0)
    s.enqueue_rounds(3000.0, 1.0, False, 3)
    auto = (s.fetch_stats(3), s.get_pose())
    s.set_mode(2)
    s.set_pose(fr["pose0"])
    with pytest.raises(vo.VoError):
        s.enqueue_rounds(3000.0, 1.0, False, 3)
    s.set_mode(1)
    s.set_pose(fr["pose0"])
    s.enqueue_rounds(3000.0, 1.0, False, 3)
    assert np.array_equal(s.get_pose(), auto[1])
    s.close()


def test_resident_is_deterministic_and_restartable(ctx):
    """two solves of the same frame on one handle (the word buffers' sequence numbers keep counting) and on a fresh
    handle give bit-identical stats and poses; rounds split over two calls equal one call"""
    fr = synth.picp_frame(n=300001, seed=21, permute=True)
    s = _solver(ctx, fr, mode=2)
    outs = []
    for rep in range(3):
        s.set_pose(fr["pose0"])
        s.enqueue_rounds(3000.0, 1.0, False, 8)
        st = s.fetch_stats(8)
        outs.append(([x.chi_inliers for x in st], [x.num_inliers for x in st], s.get_pose()))
    s.set_pose(fr["pose0"])
    s.enqueue_rounds(3000.0, 1.0, False, 3)
    s.enqueue_rounds(3000.0, 1.0, False, 5)
    st = s.fetch_stats(5)
    split_pose = s.get_pose()
    s.close()
    s2 = _solver(ctx, fr, mode=2)
    s2.enqueue_rounds(3000.0, 1.0, False, 8)
    st2 = s2.fetch_stats(8)
    outs.append(([x.chi_inliers for x in st2], [x.num_inliers for x in st2], s2.get_pose()))
    s2.close()
    for o in outs[1:]:
        assert o[0] == outs[0][0] and o[1] == outs[0][1] and np.array_equal(o[2], outs[0][2])
    assert np.array_equal(split_pose, outs[0][2])
    assert [x.num_inliers for x in st] == outs[0][1][3:]


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["stream", "resident", "stream_persistent"])
def test_out_of_range_index_on_the_device_path(ctx, mode):
    """set_correspondences_dev cannot validate synchronously: the bad index is reported by the next fetch / solve,
    and a following valid set works (the flag does not stick)."""
    import torch
    vo = product()
    fr = synth.picp_frame(n=5000, seed=3)
    bad = fr["pairs"].copy()
    bad[1234, 1] = 5000
    dw, di = torch.from_numpy(fr["world"]).cuda(), torch.from_numpy(fr["image"]).cuda()
    dbad, dgood = torch.from_numpy(bad).cuda(), torch.from_numpy(fr["pairs"]).cuda()
    torch.cuda.synchronize()
    s = ctx.picp()
    s.set_mode(mode)
    s.set_camera(fr["K"], 480, 640, fr["pose0"])
    s.set_points_dev(dw.data_ptr(), 5000, di.data_ptr(), 5000)
    s.set_correspondences_dev(dbad.data_ptr(), 5000)
    s.enqueue_rounds(3000.0, 1.0, False, 3)
    with pytest.raises(vo.VoError, match="out of range"):
        s.fetch_stats(3)
    s.set_pose(fr["pose0"])
    s.set_correspondences_dev(dgood.data_ptr(), 5000)
    s.enqueue_rounds(3000.0, 1.0, False, 3)
    st = s.fetch_stats(3)
    assert st[-1].num_inliers > 4000
    s.close()


def test_reciprocal_shortcut_exhaustive(ctx):
    """rcp.approx + one FMA Newton step == the correctly rounded reciprocal (__frcp_rn == 1.f / z) for EVERY float
    inside the gate 1e-30 <= z <= 1e30 that pair_front / picp_project apply it under: all 2^32 bit patterns."""
    in_gate, bad_packed, bad_scalar, first = ctx.selftest_reciprocal()
    lo = np.array([1e-30], np.float32).view(np.uint32)[0]
    hi = np.array([1e30], np.float32).view(np.uint32)[0]
    assert in_gate == int(hi) - int(lo) + 1  # every positive float in the gate was visited exactly once
    assert bad_packed == 0 and bad_scalar == 0, f"first mismatch at bit pattern {first - 1:#x}"


@pytest.mark.parametrize("n", [1 << 20, 10 * (1 << 20)])
@pytest.mark.parametrize("permute", [False, True], ids=["identity", "permuted"])
def test_benchmarked_sizes_match_oracle(ctx, oracle, n, permute):
    """BASELINE configs 2 and 3 AT FULL SIZE (1,048,576 and 10,485,760 correspondences, variant A identity and variant
    B permuted world indices): per-correspondence status bit-exact, counts exact, H / b / chi <= 1e-4 against the
    float64-accumulated oracle - for thr 3000 with inlier rejection and for thr 100 with kept outliers (the lambda
    branch, src/picp_solver.cpp:77-80)."""
    fr = synth.picp_frame(n=n, seed=42, permute=permute)
    s = _solver(ctx, fr)
    for thr, keep in ((3000.0, False), (100.0, True)):
        lin = s.linearize(thr, keep, want_status=True, n_pairs=n)
        ref = oracle.linearize(fr["K"], 480, 640, fr["pose0"], fr["world"], fr["image"], fr["pairs"], thr, keep,
                               accum="f64")
        _check_lin(lin, ref)
    # ten rounds (the benchmarked step): counts of every round within the knife-edge slack of the oracle's, pose 1e-5
    s.enqueue_rounds(3000.0, 1.0, False, 10)
    st = s.fetch_stats(10)
    pose = fr["pose0"].copy()
    for r in range(10):
        pose, ci, co, ni = oracle.one_round(fr["K"], 480, 640, pose, fr["world"], fr["image"], fr["pairs"], 3000.0, 1.0,
                                            False, n_threads=16)
        slack = 0 if r == 0 else max(2, int(1e-5 * n))
        assert abs(st[r].num_inliers - ni) <= slack, r
    assert np.abs(s.get_pose() - pose).max() <= POSE_TOL
    s.close()


def test_full_size_properties(ctx):
    """BASELINE config 2/3 sizes (1M and 10M correspondences): size-independent properties instead of
    the oracle: linearity of the reduction (H,b of the whole = sum over two halves), determinism,
    and convergence of 10 rounds to the generator's ground-truth pose."""
    for n in (1 << 20, 10 * (1 << 20)):
        fr = synth.picp_frame(n=n, seed=42)
        s = _solver(ctx, fr)
        whole = s.linearize(3000.0, False)
        again = s.linearize(3000.0, False)
        assert np.array_equal(whole["H"], again["H"]) and np.array_equal(whole["b"], again["b"])
        half = n // 2
        parts = []
        for lo, hi in ((0, half), (half, n)):
            s.set_correspondences(fr["pairs"][lo:hi])
            parts.append(s.linearize(3000.0, False))
        assert parts[0]["n_inliers"] + parts[1]["n_inliers"] == whole["n_inliers"]
        Hs = parts[0]["H"].astype(np.float64) + parts[1]["H"]
        assert np.abs(Hs - whole["H"]).max() <= 1e-5 * np.abs(whole["H"]).max()
        s.set_correspondences(fr["pairs"])
        s.enqueue_rounds(3000.0, 1.0, False, 10)
        stats = s.fetch_stats(10)
        assert stats[-1].num_inliers > 0.85 * n
        assert np.abs(s.get_pose() - fr["pose_gt"]).max() < 1e-3
        s.close()


@pytest.mark.parametrize("seed", [3, 4])
def test_round_kernels_randomised_sweep(seed):
    """exp/picp_rounds_stress.py: random sizes around the tile / grid / capacity boundaries, thresholds, keep_outliers,
    pinhole and general K, identity and permuted pairs, 1..12 rounds and the in-kernel convergence test - the three round
    kernels against each other (counts, chi 2e-5, pose 2e-6, same stopping round) and round 0 against the oracle"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "exp", "picp_rounds_stress.py"), str(seed), "14"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "mismatches: 0" in r.stdout
