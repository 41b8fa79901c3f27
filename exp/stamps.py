import importlib, sys, os, ctypes
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth
vo = importlib.import_module("02-visualodometry_b200")
L = ctypes.CDLL(os.environ["VO_B200_LIB"])
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = vo.Context(0, stream.cuda_stream)
fr = synth.picp_frame(n=10 * (1 << 20), seed=42)
dw = torch.from_numpy(fr["world"]).cuda(); di = torch.from_numpy(fr["image"]).cuda(); dp = torch.from_numpy(fr["pairs"]).cuda()
for C in (1024, 1 << 20, 10 << 20):
    s = ctx.picp(); s.set_camera(fr["K"], 480, 640, fr["pose0"])
    s.set_points_dev(dw.data_ptr(), len(fr["world"]), di.data_ptr(), len(fr["image"]))
    s.set_correspondences_dev(dp.data_ptr(), C)
    for it in range(3):
        s.set_pose(fr["pose0"]); s.enqueue_rounds(3000.0, 1.0, False, 2); torch.cuda.synchronize()
    prev_end = None
    rows = []
    for it in range(6):
        s.one_round(3000.0, 1.0, False)
        st = (ctypes.c_ulonglong * 8)(); L.vo_debug_stamps(st)
        t = [st[i] for i in range(5)]
        rows.append([t[1]-t[0], t[2]-t[1], t[3]-t[2], t[4]-t[3]])
    r = np.median(np.array(rows), axis=0)
    print(f"C={C:9d} ns: main loop+block reduce (block0) {r[0]:8.0f} | ->last block ticket {r[1]:8.0f} | final reduce {r[2]:7.0f} | solve+update {r[3]:7.0f}")
    s.close()
