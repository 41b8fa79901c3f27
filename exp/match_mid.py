import importlib, sys, time, torch, numpy as np, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); import synth
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
n1 = n2 = 1 << 20
A, B = synth.descriptors(n1, n2, seed=42)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
for rows in (16384, 32768, 65536, 131072, 262144):
    ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), n2, 10, pairs.data_ptr(), rows)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): n, _ = ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), n2, 10, pairs.data_ptr(), rows)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"rows {rows:8d}: {dt*1e3:7.2f} ms", flush=True)
