"""randomised parity sweep of PICP linearize: per-correspondence status bit-exact vs the oracle, H/b within 1e-4,
for random cameras (pinhole and general K), poses (any rotation), scene scales 1e-3..1e6, thresholds, keep_outliers"""
import importlib, sys, os, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth
from oracle import pyoracle as O
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bad = 0
for case in range(n_cases):
    K, rows, cols, pose, world, image, pairs, thr, keep, general, scale = synth.picp_stress_case(rng)
    n = len(pairs)
    s = ctx.picp(); s.set_camera(K, rows, cols, pose); s.set_points(world, image); s.set_correspondences(pairs)
    lin = s.linearize(thr, keep, want_status=True, n_pairs=n)
    ref = O.linearize(K, rows, cols, pose, world, image, pairs, thr, keep, accum="f64")
    st_ok = np.array_equal(lin["status"], ref["status"]) and lin["n_inliers"] == ref["n_inliers"]
    with np.errstate(all="ignore"):
        hs = np.abs(ref["H"]).max(); bs = np.abs(ref["b"]).max()
        fin = np.isfinite(hs) and np.isfinite(bs)
        if fin:
            h_ok = np.abs(lin["H"] - ref["H"]).max() <= 1e-4 * hs and np.abs(lin["b"] - ref["b"]).max() <= 1e-4 * max(bs, 1e-30)
        else:  # an inlier with non-finite terms poisons the system in the reference; it must do so here too
            h_ok = not (np.isfinite(lin["H"]).all() and np.isfinite(lin["b"]).all())
    ok = st_ok and h_ok
    bad += not ok
    cnt = np.bincount(ref["status"], minlength=3)
    print(f"case {case:2d} n {n:6d} scale {scale:g} {'general' if general else 'pinhole'} thr {thr:g} keep {int(keep)}: status {'ok' if st_ok else 'MISMATCH ' + str(int((lin['status'] != ref['status']).sum()))}  H/b {'ok' if h_ok else 'OUT'}{'' if fin else ' (non-finite ref)'}  statuses {cnt.tolist()}", flush=True)
    s.close()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
