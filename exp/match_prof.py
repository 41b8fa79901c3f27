"""one 131072 x 1048576 row-shard match (the per-GPU share of BASELINE config 4 at N = 8) for ncu"""
import importlib, sys, os, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); import synth
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
n1, n2 = (int(sys.argv[1]) if len(sys.argv) > 1 else 131072), 1 << 20
A, B = synth.descriptors(n1, n2, seed=42)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
for _ in range(2):
    n, _ = ctx.match_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, 10, pairs.data_ptr(), n1)
torch.cuda.synchronize()
print("matches", n)
