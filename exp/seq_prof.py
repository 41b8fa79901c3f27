"""592 sequences x 121 frames (one wave of 4 CTAs per SM) for ncu"""
import importlib, sys, os, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth, bench
vo = importlib.import_module("02-visualodometry_b200")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
dev = torch.device("cuda", 0)
ctx = vo.Context(0, stream.cuda_stream)
S, F, P, W = 592, 121, 128, 1024
cnt, uv, desc, ids = bench.simulate_sequences_torch(torch, dev, S, F, seed=42)
poses = torch.empty((S, F, 12), dtype=torch.float32, device=dev); wxyz = torch.empty((S, W, 3), dtype=torch.float32, device=dev)
wid = torch.empty((S, W), dtype=torch.int32, device=dev); wcnt = torch.empty(S, dtype=torch.int32, device=dev); status = torch.empty(S, dtype=torch.int32, device=dev)
params = vo.seq_params(synth.K_REF)
ctx.seq_batch_run_dev(params, S, F, P, W, cnt.data_ptr(), uv.data_ptr(), desc.data_ptr(), ids.data_ptr(), poses.data_ptr(), wxyz.data_ptr(), wid.data_ptr(), wcnt.data_ptr(), None, None, status.data_ptr())
torch.cuda.synchronize()
print("ok", int((status == 0).sum()))
