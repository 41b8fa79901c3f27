import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
import test_gpu_picp as T
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
probe = ctx.picp(); cap = probe.resident_capacity; probe.close()
for n in T._sizes_around_resident_geometry(cap):
    permute = n % 2 == 1
    thr, keep = ((3000.0, False), (100.0, True))[(n // 3) % 2]
    fr = synth.picp_frame(n=n, seed=7 + n, permute=permute)
    res = {}
    for mode in (1, 2, 3):
        s = T._solver(ctx, fr, mode=mode)
        s.enqueue_rounds(thr, 1.0, keep, 6)
        res[mode] = (s.fetch_stats(6), s.get_pose())
        s.close()
    for r in range(6):
        a, b, c = res[1][0][r], res[2][0][r], res[3][0][r]
        rel = abs(a.chi_inliers - b.chi_inliers) / max(a.chi_inliers, 1.0)
        rel3 = abs(a.chi_inliers - c.chi_inliers) / max(a.chi_inliers, 1.0)
        flag = " <<<" if rel > 1e-5 or rel3 > 1e-5 else ""
        print(n, keep, r, a.num_inliers, b.num_inliers, c.num_inliers, a.chi_inliers, b.chi_inliers, c.chi_inliers, "%.2e %.2e" % (rel, rel3), flag)
    print(n, "pose diff", np.abs(res[1][1] - res[2][1]).max(), np.abs(res[1][1] - res[3][1]).max())
