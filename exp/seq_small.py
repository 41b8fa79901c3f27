"""a few short sequences for compute-sanitizer"""
import importlib, sys, os, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth, bench
vo = importlib.import_module("02-visualodometry_b200")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
dev = torch.device("cuda", 0)
ctx = vo.Context(0, stream.cuda_stream)
S, F, P, W = 3, int(sys.argv[1]) if len(sys.argv) > 1 else 6, 128, 1024
cnt, uv, desc, ids = bench.simulate_sequences_torch(torch, dev, S, F, seed=42)
poses = torch.zeros((S, F, 12), dtype=torch.float32, device=dev); wxyz = torch.zeros((S, W, 3), dtype=torch.float32, device=dev)
wid = torch.zeros((S, W), dtype=torch.int32, device=dev); wcnt = torch.zeros(S, dtype=torch.int32, device=dev); status = torch.zeros(S, dtype=torch.int32, device=dev)
rounds = torch.zeros((S, F), dtype=torch.int32, device=dev); inl = torch.zeros((S, F, 2), dtype=torch.int32, device=dev)
params = vo.seq_params(synth.K_REF)
ctx.seq_batch_run_dev(params, S, F, P, W, cnt.data_ptr(), uv.data_ptr(), desc.data_ptr(), ids.data_ptr(), poses.data_ptr(), wxyz.data_ptr(), wid.data_ptr(), wcnt.data_ptr(), rounds.data_ptr(), inl.data_ptr(), status.data_ptr())
torch.cuda.synchronize()
print("status", status.tolist(), "wcnt", wcnt.tolist())
print(poses[0, :F, 3].tolist())
print("rounds", rounds[0].tolist())
print("inl", inl[0].tolist())
print("pose1", [round(x, 5) for x in poses[0, 1].tolist()])
print("wxyz", [[round(x, 4) for x in r] for r in wxyz[0, :4].tolist()])
print("wid", wid[0, :12].tolist())
