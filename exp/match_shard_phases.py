"""one rank's share of the 1M x 1M sharded matching (shard 3 of 8, no communicator: the exchange is skipped) on one GPU:
wall time per call, for `ncu --metrics gpu__time_duration.sum` to list the kernels of ONE call (VO_PHASES_ONCE=1)"""
import importlib, os, sys, time
import numpy as np
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
n1 = n2 = 1 << 20
A, B = synth.descriptors(n1, n2, seed=42)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
midx = torch.empty(n1, dtype=torch.int32, device="cuda")
pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
once = os.environ.get("VO_PHASES_ONCE") == "1"
for n_shards in ((8,) if once else (1, 2, 4, 8)):
    shard = min(3, n_shards - 1)
    run = lambda: ctx.match_sharded_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, 10, shard, n_shards, midx.data_ptr(), pairs.data_ptr(), n1)
    if once:
        run(); torch.cuda.synchronize()
        break
    run(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        n = run()
    torch.cuda.synchronize()
    print(f"shard {shard} of {n_shards}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call, {n} pairs from this shard's rows")
