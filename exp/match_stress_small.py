"""the same adversarial descriptor sets on the generic split scan and the Morton-ordered packed scan (n2 < 8192)"""
import importlib, sys, os, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth
from oracle import pyoracle as O
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    kind = str(rng.choice(["uniform", "clustered", "lattice", "lowrank", "cauchy"]))
    n1 = int(rng.choice([1, 33, 500, 4097, 9000, 20000])); n2 = int(rng.choice([1, 2, 7, 490, 4099, 8191]))
    A, B = synth.stress_descriptors(rng, kind, n1, n2)
    thr = (0.2, 0.8) if rng.random() < 0.5 else (float(np.float32(np.median(np.abs(A)) ** 2 * 4 + 1e-30)), 1.5)
    rp, _, rbest, rsecond, ridx = O.match(A, B, dist_thr=thr[0], ratio_thr=thr[1], want_rows=True, n_threads=16)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    best = torch.empty(n1, dtype=torch.float32, device="cuda"); second = torch.empty_like(best)
    idx = torch.empty(n1, dtype=torch.int32, device="cuda"); pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
    n, _ = ctx.match_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, 10, pairs.data_ptr(), n1, dist_thr=thr[0], ratio_thr=thr[1],
                         d_best=best.data_ptr(), d_second=second.data_ptr(), d_idx=idx.data_ptr())
    torch.cuda.synchronize()
    ok = (np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(best.cpu().numpy().view(np.uint32), rbest.view(np.uint32))
          and np.array_equal(second.cpu().numpy().view(np.uint32), rsecond.view(np.uint32)) and np.array_equal(pairs[:n].cpu().numpy(), rp))
    bad += not ok
    print(f"case {case:2d} {kind:9s} {n1:6d} x {n2:5d}: {'ok' if ok else 'MISMATCH'}  {n} pairs", flush=True)
print("mismatches:", bad)
