"""times vo_match_dev on BASELINE config 4 (1M x 1M) and on a 131072-row shard with the library named by VO_B200_LIB"""
import importlib, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
A, B = synth.descriptors(1 << 20, 1 << 20, seed=42)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
pairs = torch.empty((1 << 20, 2), dtype=torch.int32, device="cuda")
for rows in (1 << 20, 131072):
    ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), 1 << 20, 10, pairs.data_ptr(), 1 << 20)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        t0 = time.perf_counter()
        n, _ = ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), 1 << 20, 10, pairs.data_ptr(), 1 << 20)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(os.path.basename(vo.LIB_PATH), "rows", rows, "median ms %.3f" % (1e3 * sorted(ts)[3]), "matches", n)
