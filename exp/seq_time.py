import importlib, sys, os, torch, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth, bench
vo = importlib.import_module("02-visualodometry_b200")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
dev = torch.device("cuda", 0)
ctx = vo.Context(0, stream.cuda_stream)
for total in (4096, 512):
    class A: pass
    r = bench.bench_sequences.__wrapped__ if hasattr(bench.bench_sequences, "__wrapped__") else None
    S, F, P, W = total, 121, 128, 1024
    cnt, uv, desc, ids = bench.simulate_sequences_torch(torch, dev, S, F, seed=42)
    poses = torch.empty((S, F, 12), dtype=torch.float32, device=dev); wxyz = torch.empty((S, W, 3), dtype=torch.float32, device=dev)
    wid = torch.empty((S, W), dtype=torch.int32, device=dev); wcnt = torch.empty(S, dtype=torch.int32, device=dev); status = torch.empty(S, dtype=torch.int32, device=dev)
    params = vo.seq_params(synth.K_REF)
    run = lambda: ctx.seq_batch_run_dev(params, S, F, P, W, cnt.data_ptr(), uv.data_ptr(), desc.data_ptr(), ids.data_ptr(), poses.data_ptr(), wxyz.data_ptr(), wid.data_ptr(), wcnt.data_ptr(), None, None, status.data_ptr())
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); run(); run(); b.record(stream); torch.cuda.synchronize()
    print(f"{total} sequences: {a.elapsed_time(b)/2:.1f} ms  ok {(status==0).sum().item()}  checksum {poses.double().sum().item():.6f}")
