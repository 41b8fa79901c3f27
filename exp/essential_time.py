"""latency of vo_essential_recover (RANSAC five-point on the GPU) on the bundled dataset's first frame pair (115 clean
matches: 1 iteration) and on a 400-match problem with 30 % gross outliers, beside the oracle port on the host"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import replay, synth
from oracle import pyoracle as O
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
ds = dict(np.load(os.path.join(ROOT, "tests", "golden", "dataset.npz")))
f0, f1 = replay.frame(ds, 0), replay.frame(ds, 1)
m, _ = ctx.match(f0["desc"], f1["desc"])
cases = {"dataset 0/1 (115 matches, clean)": (f0["uv"][m[:, 0]], f1["uv"][m[:, 1]])}
rng = np.random.default_rng(5)
X = np.stack([rng.normal(0, 2, 400), rng.normal(0, 1.5, 400), rng.uniform(3, 15, 400)], 1)
rel = synth.euler_pose(np.array([0.3, -0.1, 0.8, 0.04, -0.06, 0.03]))
def proj(T):
    c = (X - T[:, 3]) @ T[:, :3]; q = c @ synth.K_REF.astype(np.float64).T
    return (q[:, :2] / q[:, 2:3]).astype(np.float32)
x1, x2 = proj(np.eye(4)[:3]), proj(rel)
bad = rng.random(400) < 0.3
x2[bad] = rng.uniform(0, 480, (int(bad.sum()), 2)).astype(np.float32)
cases["400 matches, 30 % gross outliers"] = (x1, x2)
for name, (a, b) in cases.items():
    for method in ("ransac", "8pt"):
        ctx.essential_recover(synth.K_REF, a, b, method=method)
        t0 = time.perf_counter()
        for _ in range(5):
            out = ctx.essential_recover(synth.K_REF, a, b, method=method, full=True)
        dt = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(5):
            O.essential_recover(synth.K_REF, a, b, method=method)
        dto = (time.perf_counter() - t0) / 5
        print(f"{name:36s} {method:7s} GPU {dt * 1e3:7.3f} ms (iterations {out[7]}, inliers {out[6]})   oracle on the host {dto * 1e3:7.3f} ms")
