"""profiling workload for seq_pipeline_kernel: 888 simulated 121-frame sequences (one full wave: 148 SMs x 6 CTAs)"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
import bench
vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
dev = torch.device("cuda", 0)
S, F, P, W = 888, 121, 128, 1024
cnt, uv, desc, ids = bench.simulate_sequences_torch(torch, dev, S, F, seed=42)
poses = torch.empty((S, F, 12), dtype=torch.float32, device=dev); wxyz = torch.empty((S, W, 3), dtype=torch.float32, device=dev)
wid = torch.empty((S, W), dtype=torch.int32, device=dev); wcnt = torch.empty(S, dtype=torch.int32, device=dev)
status = torch.empty(S, dtype=torch.int32, device=dev)
for _ in range(2):
    ctx.seq_batch_run_dev(vo.seq_params(synth.K_REF), S, F, P, W, cnt.data_ptr(), uv.data_ptr(), desc.data_ptr(), ids.data_ptr(),
                          poses.data_ptr(), wxyz.data_ptr(), wid.data_ptr(), wcnt.data_ptr(), None, None, status.data_ptr())
    ctx.sync()
print("ok", int((status == 0).sum()), "of", S)
