"""Profiling workload for the two PICP kernels (ncu --set full -k regex:picp_): 3 solves of BASELINE config 2's 1M
frame on the resident kernel (one launch = 10 rounds) and 1 solve of the 10M frame on the streaming kernel."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
for n, reps in ((1 << 20, 3), (10 * (1 << 20), 1)):
    fr = synth.picp_frame(n=n, seed=42)
    dw, di, dp = (torch.from_numpy(fr[k]).cuda() for k in ("world", "image", "pairs"))
    s = ctx.picp()
    s.set_camera(fr["K"], 480, 640, fr["pose0"])
    s.set_points_dev(dw.data_ptr(), n, di.data_ptr(), n)
    for _ in range(reps):
        s.set_pose(fr["pose0"])
        s.set_correspondences_dev(dp.data_ptr(), n)
        s.enqueue_rounds(3000.0, 1.0, False, 10)
        st = s.fetch_stats(10)
    print(n, st[-1].num_inliers, np.abs(s.get_pose() - fr["pose_gt"]).max())
    s.close()
