"""how much of the column set the indexed matcher touches (needs exp/dbg/libvo_dbg.so built with -DVO_MATCH_COUNTERS)"""
import importlib, sys, os, ctypes, time, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["VO_B200_LIB"] = os.path.join(R, "exp/dbg/libvo_dbg.so")
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); import synth
vo = importlib.import_module("02-visualodometry_b200")
lib = ctypes.CDLL(os.environ["VO_B200_LIB"])
ctx = vo.Context(0)
n1 = n2 = 1 << 20
A, B = synth.descriptors(n1, n2, seed=42)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
c = (ctypes.c_ulonglong * 4)()
for rows in (8192, 131072, n1):
    ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), n2, 10, pairs.data_ptr(), rows)
    torch.cuda.synchronize(); lib.vo_debug_match_counters(c, 1)
    t0 = time.perf_counter()
    ctx.match_dev(dA.data_ptr(), rows, dB.data_ptr(), n2, 10, pairs.data_ptr(), rows)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    lib.vo_debug_match_counters(c, 1)
    ctas = (rows + 31) // 32
    print(f"rows {rows}: {dt*1e3:.1f} ms; per CTA: super {c[0]/ctas:.1f}/512, tiles {c[1]/ctas:.1f}/8192, "
          f"pairs bounded {c[3]/ctas:.0f}, finished {c[2]/ctas:.0f} ({100*c[2]/max(1,c[3]):.1f}%)")
