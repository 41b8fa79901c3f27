"""randomised sweep of triangulation, essential + recoverPose, projectPoints and the anti-join against the oracle"""
import importlib, sys, os, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth
from oracle import pyoracle as O
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    n = int(rng.choice([8, 9, 50, 490, 5000, 100000]))
    K = np.array([[rng.uniform(100, 800), 0, rng.uniform(200, 500)], [0, rng.uniform(100, 800), rng.uniform(150, 400)], [0, 0, 1]], np.float32)
    scale = float(rng.choice([0.01, 1.0, 1.0, 100.0]))
    base = rng.normal(0, 1, 3); base *= scale * rng.choice([1e-4, 0.05, 0.5, 3.0]) / np.linalg.norm(base)   # incl. near-zero baseline
    ang = rng.normal(0, 1, 3) * rng.choice([0.0, 0.01, 0.3])
    T1 = synth.euler_pose(np.array([*(rng.normal(0, 1, 3) * scale), *(rng.normal(0, 0.3, 3))]))
    rel = synth.euler_pose(np.array([*base, *ang]))
    T2 = np.zeros((3, 4)); T2[:, :3] = T1[:, :3] @ rel[:, :3]; T2[:, 3] = T1[:, :3] @ rel[:, 3] + T1[:, 3]   # camera-in-world poses
    def proj(T, X):
        c = (X - T[:, 3]) @ T[:, :3]
        q = c @ K.astype(np.float64).T
        return q[:, :2] / q[:, 2:3], c[:, 2]
    Xc = np.stack([rng.normal(0, 1.5, n), rng.normal(0, 1.0, n), rng.uniform(0.5, 30, n)], 1) * scale
    far = rng.random(n) < 0.02
    Xc[far, 2] *= 1e6                                                  # points at (almost) infinity
    X = Xc @ T1[:, :3].T + T1[:, 3]
    x1, z1 = proj(T1, X); x2, z2 = proj(T2, X)
    noise = rng.choice([0.0, 0.3, 2.0])
    x1 = (x1 + rng.normal(0, noise, x1.shape)).astype(np.float32); x2 = (x2 + rng.normal(0, noise, x2.shape)).astype(np.float32)
    same = rng.random(n) < 0.01
    x2[same] = x1[same]                                                # zero parallax
    T1f, T2f = T1.astype(np.float32), T2.astype(np.float32)
    ok = True; notes = []
    # triangulation
    g = ctx.triangulate(K, T1f, T2f, x1, x2); r = O.triangulate(K, T1f, T2f, x1, x2)
    with np.errstate(all="ignore"):
        both = np.isfinite(g).all(1) & np.isfinite(r).all(1)
        ext = np.abs(r[both]).max() if both.any() else 1.0
        # a point whose homogeneous coordinate is ~0 is ill-conditioned in both; compare the well-conditioned ones
        well = both & (np.abs(r).max(1) < 1e4 * scale)
        dt = np.abs(g[well] - r[well]).max() / max(np.abs(r[well]).max(), 1e-30) if well.any() else 0.0
    fin_same = np.array_equal(np.isfinite(g).all(1), np.isfinite(r).all(1))
    if not (dt <= 1e-4 and fin_same): ok = False; notes.append(f"tri {dt:.2e} finite-pattern {fin_same}")
    # essential + recoverPose
    try:
        Eg, Rg, tg, mg, gg = ctx.essential_recover(K, x1, x2)
        Er, Rr, tr, mr, gr = O.essential_recover(K, x1, x2)
        de = min(np.abs(Eg - Er).max(), np.abs(Eg + Er).max()); dr = np.abs(Rg - Rr).max(); dtt = np.abs(tg - tr).max()
        mm = int((mg != mr).sum())
        if int(gg) == 0 and int(gr) == 0:
            # zero baseline without noise: the epipolar system has a multi-dimensional null space, no candidate passes
            # the cheirality vote in either implementation and the returned direction of t is arbitrary
            notes.append("degenerate (no cheirality-good point in either)")
        elif not (dr <= 1e-6 and dtt <= 1e-6 and mm <= max(1, n // 2000) and abs(int(gg) - int(gr)) <= max(1, n // 2000)):
            ok = False; notes.append(f"ess dE {de:.1e} dR {dr:.1e} dt {dtt:.1e} mask diff {mm} good {gg}/{gr}")
    except vo.VoError as e:
        notes.append(f"ess error {e}")
    # projectPoints (bit-exact) and anti-join (bit-exact)
    pose_wic = O.pose_inverse(T2f)
    pg, ig = ctx.project_points(K, 480, 640, pose_wic, X.astype(np.float32)); pr_, ir = O.project_points(K, 480, 640, pose_wic, X.astype(np.float32))
    if not (np.array_equal(pg.view(np.uint32), pr_.view(np.uint32)) and ig == ir): ok = False; notes.append("project")
    a = rng.integers(0, 3 * n, n).astype(np.int32); b = rng.integers(0, 3 * n, int(rng.integers(1, 2 * n))).astype(np.int32)
    if not np.array_equal(ctx.anti_join(a, b), O.anti_join(a, b)): ok = False; notes.append("antijoin")
    bad += not ok
    print(f"case {case:2d} n {n:6d} scale {scale:g} |base| {np.linalg.norm(base):.2e} noise {noise}: {'ok' if ok else 'MISMATCH'} {'; '.join(notes)}", flush=True)
print("mismatches:", bad)
