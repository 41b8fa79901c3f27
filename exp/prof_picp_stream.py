"""ncu target: ONE 10-round solve of the 10,485,760-correspondence frame on the persistent streaming kernel
(ncu --set full -k regex:picp_stream -c 1)"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
n = 10 * (1 << 20)
fr = synth.picp_frame(n=n, seed=42)
dw, di, dp = (torch.from_numpy(fr[k]).cuda() for k in ("world", "image", "pairs"))
s = ctx.picp()
s.set_camera(fr["K"], 480, 640, fr["pose0"])
s.set_points_dev(dw.data_ptr(), n, di.data_ptr(), n)
s.set_pose(fr["pose0"])
s.set_correspondences_dev(dp.data_ptr(), n)
s.enqueue_rounds(3000.0, 1.0, False, 10)
st = s.fetch_stats(10)
print(n, st[-1].num_inliers, np.abs(s.get_pose() - fr["pose_gt"]).max())
s.close()
