"""time per Gauss-Newton round vs correspondence count: separates the fixed per-launch cost from the streaming slope"""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import synth
vo = importlib.import_module("02-visualodometry_b200")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = vo.Context(0, stream.cuda_stream)
fr = synth.picp_frame(n=10 * (1 << 20), seed=42)
dw = torch.from_numpy(fr["world"]).cuda(); di = torch.from_numpy(fr["image"]).cuda(); dp = torch.from_numpy(fr["pairs"]).cuda()
for C in (1024, 65536, 1 << 20, 2 << 20, 4 << 20, 10 << 20):
    s = ctx.picp(); s.set_camera(fr["K"], 480, 640, fr["pose0"])
    s.set_points_dev(dw.data_ptr(), len(fr["world"]), di.data_ptr(), len(fr["image"]))
    s.set_correspondences_dev(dp.data_ptr(), C)
    ts = []
    for it in range(8):
        s.set_pose(fr["pose0"])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); s.enqueue_rounds(3000.0, 1.0, False, 10); b.record(stream); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 100)  # us per round
    ts.sort()
    print(f"C={C:9d}  {ts[len(ts)//2]:8.2f} us/round")
    s.close()
