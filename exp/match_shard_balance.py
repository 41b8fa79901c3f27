"""per-shard time of the 8-way (and 4-way) curve-sharded matching on one GPU (no communicator): load balance"""
import importlib, os, sys, time
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
for n_shards in (8, 4):
    ts = []
    for shard in range(n_shards):
        run = lambda: ctx.match_sharded_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, 10, shard, n_shards, midx.data_ptr(), pairs.data_ptr(), n1)
        run(); torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / 3)
        ts.append(best * 1e3)
    print(f"{n_shards} shards: " + " ".join(f"{t:.2f}" for t in ts) + f"   max {max(ts):.2f} mean {sum(ts)/len(ts):.2f} ms")
