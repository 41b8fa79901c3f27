import importlib, sys, time, torch, numpy as np, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); import synth
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
n1 = n2 = 1 << 20
A, B = synth.descriptors(n1, n2, seed=42)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
for rows in (8192, 32768, 131072, 524288, n1):
    ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), n2, 10, pairs.data_ptr(), rows)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): n, _ = ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), n2, 10, pairs.data_ptr(), rows)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"rows {rows:8d}: {dt*1e3:7.2f} ms  {rows*n2/dt/1e12:.3f} Tpairs/s")
