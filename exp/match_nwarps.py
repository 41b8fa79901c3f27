"""warps per 32-row group of the filtered scan (VO_MATCH_NWARPS, read per call) x shard count: best of 3 x 4 calls"""
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
for _ in range(3):
    ctx.match_sharded_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, 10, 0, 1, midx.data_ptr(), pairs.data_ptr(), n1)
for n_shards in (1, 2, 4, 8):
    shard = min(3, n_shards - 1)
    line = []
    for nw in (1, 2, 4, 8):
        if n_shards == 1 and nw > 2:
            continue
        os.environ["VO_MATCH_NWARPS"] = str(nw)
        run = lambda: ctx.match_sharded_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, 10, shard, n_shards, midx.data_ptr(), pairs.data_ptr(), n1)
        run(); torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for _ in range(4):
                run()
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / 4)
        line.append(f"{nw} warps {best * 1e3:6.2f} ms")
    print(f"shard {shard} of {n_shards}: " + "   ".join(line))
# fewer rows against all 1M columns (one call = index build + scan + compaction)
for rows in (2048, 8192, 16384, 32768, 65536):
    line = []
    for nw in (1, 2, 4, 8, 16):
        os.environ["VO_MATCH_NWARPS"] = str(nw)
        run = lambda: ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), n2, 10, pairs.data_ptr(), rows)
        run(); torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for _ in range(4):
                run()
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / 4)
        line.append(f"{nw} warps {best * 1e3:6.2f} ms")
    os.environ.pop("VO_MATCH_NWARPS")
    run(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        run()
    torch.cuda.synchronize()
    print(f"rows {rows}: " + "   ".join(line) + f"   | default rule {(time.perf_counter() - t0) / 4 * 1e3:6.2f} ms")
