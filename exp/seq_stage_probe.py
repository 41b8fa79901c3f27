"""per-stage clock split of seq_pipeline_kernel (library built with -DVO_SEQ_STAGE_CLOCKS: exp/build_variant.sh seqclk):
1024 simulated 121-frame sequences; thread 0 of every CTA accumulates clock64() per stage of the frame loop."""
import ctypes, importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
import bench
vo = importlib.import_module("02-visualodometry_b200")
lib = ctypes.CDLL(vo.LIB_PATH)
ctx = vo.Context(0)
dev = torch.device("cuda", 0)
S, F, P, W = 1024, 121, 128, 1024
cnt, uv, desc, ids = bench.simulate_sequences_torch(torch, dev, S, F, seed=42)
poses = torch.empty((S, F, 12), dtype=torch.float32, device=dev); wxyz = torch.empty((S, W, 3), dtype=torch.float32, device=dev)
wid = torch.empty((S, W), dtype=torch.int32, device=dev); wcnt = torch.empty(S, dtype=torch.int32, device=dev)
status = torch.empty(S, dtype=torch.int32, device=dev); rounds = torch.empty((S, F), dtype=torch.int32, device=dev)
params = vo.seq_params(synth.K_REF)
def run():
    ctx.seq_batch_run_dev(params, S, F, P, W, cnt.data_ptr(), uv.data_ptr(), desc.data_ptr(), ids.data_ptr(), poses.data_ptr(),
                          wxyz.data_ptr(), wid.data_ptr(), wcnt.data_ptr(), rounds.data_ptr(), None, status.data_ptr())
    ctx.sync()
run()
out = (ctypes.c_ulonglong * 16)()
lib.vo_debug_seq_stage_cycles(out, 1)
run()
lib.vo_debug_seq_stage_cycles(out, 1)
names = ["frame loads", "match vs map", "PICP rounds", "pose + match vs previous frame + anti-join", "triangulate + append", "initialisation (frames 0/1)"]
tot = float(sum(out[:6]))
print("ok", int((status == 0).sum()), "of", S, "mean rounds/frame", float(rounds[:, 1:].float().mean()))
for n, c in zip(names, out[:6]):
    print(f"{n:45s} {100.0 * c / tot:5.1f} %   {c / S / 1e6:8.3f} Mcycles per sequence")

if sum(out[8:14]):
    rn = ["linearize + warp reduction", "barrier 1", "cross-warp sums + gather to lane 0", "6x6 solve + verdict", "pose update", "barrier 2"]
    n_rounds = float(rounds[:, 1:].sum())
    rt = float(sum(out[8:14]))
    print("Gauss-Newton round split (thread 0), cycles per round:")
    for n, c in zip(rn, out[8:14]):
        print(f"  {n:38s} {c / n_rounds:8.0f}  {100.0 * c / rt:5.1f} %")
