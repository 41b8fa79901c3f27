"""randomised parity sweep of the indexed / tensor-core matcher path against the oracle (values, indices, pairs)"""
import importlib, sys, os, time, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from oracle import pyoracle as O
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30

import synth

bad = 0
for case in range(n_cases):
    kind, A, B, dist_thr, ratio_thr = synth.stress_case(rng)
    n1, n2 = len(A), len(B)
    rp, _, rbest, rsecond, ridx = O.match(A, B, dist_thr=float(dist_thr), ratio_thr=float(ratio_thr), want_rows=True, n_threads=16)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    best = torch.empty(n1, dtype=torch.float32, device="cuda"); second = torch.empty_like(best)
    idx = torch.empty(n1, dtype=torch.int32, device="cuda"); pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
    t0 = time.perf_counter()
    n, _ = ctx.match_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, 10, pairs.data_ptr(), n1, dist_thr=float(dist_thr), ratio_thr=float(ratio_thr),
                         d_best=best.data_ptr(), d_second=second.data_ptr(), d_idx=idx.data_ptr())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ok = (np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(best.cpu().numpy().view(np.uint32), rbest.view(np.uint32))
          and np.array_equal(second.cpu().numpy().view(np.uint32), rsecond.view(np.uint32)) and np.array_equal(pairs[:n].cpu().numpy(), rp))
    bad += not ok
    if not ok:
        gi, gb, gs = idx.cpu().numpy(), best.cpu().numpy(), second.cpu().numpy()
        mi = np.nonzero(gi != ridx)[0]; mb = np.nonzero(gb.view(np.uint32) != rbest.view(np.uint32))[0]; ms = np.nonzero(gs.view(np.uint32) != rsecond.view(np.uint32))[0]
        print(f"   idx mismatches {len(mi)}, best {len(mb)}, second {len(ms)}; max|A| {np.abs(A).max():.3g} max|B| {np.abs(B).max():.3g} min nonzero |B| {np.abs(B[B != 0]).min():.3g}")
        for r in list(mb[:3]) + list(ms[:3]) + list(mi[:3]):
            print(f"   row {r}: gpu best {gb[r]:.9g} second {gs[r]:.9g} idx {gi[r]} | ref best {rbest[r]:.9g} second {rsecond[r]:.9g} idx {ridx[r]} | |a|^2 {float((A[r].astype(np.float64)**2).sum()):.6g} |b_ref|^2 {float((B[ridx[r]].astype(np.float64)**2).sum()) if ridx[r] >= 0 else -1:.6g}")
    print(f"case {case:2d} {kind:9s} {n1:6d} x {n2:7d} thr {float(dist_thr):.3g}/{ratio_thr}: {'ok ' if ok else 'MISMATCH'} {dt*1e3:8.2f} ms  {n} pairs", flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
