"""randomised parity sweep of the three PICP round kernels (one launch per round / persistent streaming / shared-memory
resident) against each other and against the oracle: random sizes (around the tile, grid and capacity boundaries),
thresholds, keep_outliers, pinhole and general K, identity and permuted correspondences, 1..12 rounds, in-kernel
convergence test"""
import importlib, os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from oracle import pyoracle as O
import synth
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sizes = [1, 2, 7, 120, 1407, 1408, 1409, 1536, 1537, 5633, 40000, 208385, 208500, 600001, 1310720, 1667072, 1667073, 2500000]
bad = 0
for case in range(n_cases):
    n = int(rng.choice(sizes))
    permute = bool(rng.random() < 0.5)
    thr, keep = [(3000.0, False), (100.0, True), (1000.0, False), (25.0, False)][int(rng.integers(0, 4))]
    rounds = int(rng.integers(1, 13))
    fr = synth.picp_frame(n=n, seed=int(rng.integers(0, 1 << 30)), permute=permute, outlier_frac=float(rng.choice([0.0, 0.1, 0.4])),
                          invalid_frac=float(rng.choice([0.0, 0.02, 0.3])))
    if rng.random() < 0.3:
        fr["K"] = np.array([[181.5, 0.7, 318.2], [0.01, 179.3, 241.1], [1e-4, -2e-4, 1.01]], np.float32)
    out = {}
    for mode in (1, 2, 3):
        s = ctx.picp(); s.set_mode(mode)
        s.set_camera(fr["K"], 480, 640, fr["pose0"]); s.set_points(fr["world"], fr["image"]); s.set_correspondences(fr["pairs"])
        try:
            s.enqueue_rounds(thr, 1.0, keep, rounds)
            st = s.fetch_stats(rounds)
            pose = s.get_pose()
            s.set_pose(fr["pose0"])
            done, last = s.solve(thr, 1.0, keep, max_rounds=30, rel_tol=1e-3)
            out[mode] = (st, pose, done, s.get_pose())
        except vo.VoError as e:
            out[mode] = None if (mode == 2 and "does not fit" in str(e)) else ("error", str(e))
        s.close()
    ref_pose = fr["pose0"].copy()
    ref = []
    for r in range(min(rounds, 3)):
        ref_pose, ci, co, ni = O.one_round(fr["K"], 480, 640, ref_pose, fr["world"], fr["image"], fr["pairs"], thr, 1.0, keep, n_threads=8)
        ref.append((ni, ci))
    ok = True
    why = ""
    base = out[1]
    for mode in (2, 3):
        o = out[mode]
        if o is None:
            ok = ok and n > 1667072; why += "" if n > 1667072 else " resident refused a set that fits;"
            continue
        if isinstance(o, tuple) and o[0] == "error":
            ok = False; why += f" mode {mode}: {o[1]};"; continue
        slack = max(2, int(1e-5 * n))
        for r in range(rounds):
            if abs(o[0][r].num_inliers - base[0][r].num_inliers) > (0 if r == 0 else slack): ok = False; why += f" m{mode} r{r} inl {o[0][r].num_inliers} vs {base[0][r].num_inliers};"
            # from round 1 on the poses of two kernels agree to ~1e-7 only: a correspondence whose chi sits on the threshold may
            # flip, and every flip moves chi_inliers by up to the threshold itself
            chi_tol = 2e-5 * max(base[0][r].chi_inliers, 1.0) + (0 if r == 0 else slack * thr)
            if abs(o[0][r].chi_inliers - base[0][r].chi_inliers) > chi_tol: ok = False; why += f" m{mode} r{r} chi;"
        if np.abs(o[1] - base[1]).max() > 2e-6: ok = False; why += f" m{mode} pose {np.abs(o[1] - base[1]).max():.2e};"
        if abs(o[2] - base[2]) > 1: ok = False; why += f" m{mode} solve rounds {o[2]} vs {base[2]};"
        if o[2] == base[2] and np.abs(o[3] - base[3]).max() > 2e-6: ok = False; why += f" m{mode} solve pose;"
    if base[0][0].num_inliers != ref[0][0]: ok = False; why += f" oracle inliers {ref[0][0]} vs {base[0][0].num_inliers};"
    bad += not ok
    print(f"case {case:2d} n {n:8d} perm {int(permute)} thr {thr:g} keep {int(keep)} rounds {rounds:2d} generalK {int(not np.array_equal(fr['K'], synth.K_REF))}: {'ok' if ok else 'MISMATCH' + why}", flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
