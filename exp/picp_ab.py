"""us per Gauss-Newton round of the persistent streaming kernel (10,485,760 correspondences) and of the resident kernel
(1,048,576 and 1,310,720) with the library named by VO_B200_LIB (kernel A/B experiments)"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
vo = importlib.import_module("02-visualodometry_b200")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = vo.Context(0, stream.cuda_stream)
for n in (10 * (1 << 20), 1 << 20, 1310720):
    fr = synth.picp_frame(n=n, seed=42)
    dw, di, dp = (torch.from_numpy(fr[k]).cuda() for k in ("world", "image", "pairs"))
    s = ctx.picp()
    s.set_camera(fr["K"], 480, 640, fr["pose0"])
    s.set_points_dev(dw.data_ptr(), n, di.data_ptr(), n)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    for rounds in (2, 10):
        ts = []
        for it in range(12):
            s.set_pose(fr["pose0"])
            s.set_correspondences_dev(dp.data_ptr(), n)
            if n > s.resident_capacity:
                s.pack()
            a.record(stream)
            s.enqueue_rounds(3000.0, 1.0, False, rounds)
            b.record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        res[rounds] = sorted(ts)[len(ts) // 2]
    print(os.path.basename(vo.LIB_PATH), n, "us per round (marginal): %.2f" % ((res[10] - res[2]) / 8 * 1e3), " 10 rounds: %.1f us" % (res[10] * 1e3),
          "pose err %.1e" % np.abs(s.get_pose() - fr["pose_gt"]).max())
    s.close()
