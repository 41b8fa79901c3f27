import importlib, sys, os, torch, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import synth, bench, replay, backends, simulator
vo = importlib.import_module("02-visualodometry_b200")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
dev = torch.device("cuda", 0)
ctx = vo.Context(0, stream.cuda_stream)
S, F, P, W = 4096, 121, 128, 1024
cnt, uv, desc, ids = bench.simulate_sequences_torch(torch, dev, S, F, seed=42)
out = ctx.seq_batch_run(vo.seq_params(synth.K_REF), cnt.cpu().numpy(), uv.cpu().numpy(), desc.cpu().numpy(), ids.cpu().numpy())
bad = np.nonzero(~np.isfinite(out["poses"]).all(axis=(1, 2, 3)))[0]
print("non-finite sequences:", len(bad), bad[:10], "min cnt", cnt.min().item())
big = np.nonzero(np.abs(np.nan_to_num(out["poses"])).max(axis=(1,2,3)) > 1e4)[0]
print("huge-pose sequences:", len(big), big[:10])
for s in list(bad[:2]):
    first_bad = np.nonzero(~np.isfinite(out["poses"][s]).all(axis=(1, 2)))[0][0]
    print("seq", s, "first bad frame", first_bad, "cnt around", cnt[s, max(0,first_bad-2):first_bad+2].tolist(), "inliers", out["inliers"][s, max(0,first_bad-2):first_bad+2].tolist(), "rounds", out["rounds"][s, max(0,first_bad-2):first_bad+2].tolist())
    batch = dict(cnt=cnt[s:s+1].cpu().numpy(), uv=uv[s:s+1].cpu().numpy(), desc=desc[s:s+1].cpu().numpy(), id_real=ids[s:s+1].cpu().numpy(), gt_pose=np.zeros((1, F, 3), np.float32))
    ds = simulator.as_dataset(batch, 0)
    ref = replay.run_icp_test(ds, backends.OracleBackend(), n_meas=F)
    rb = np.nonzero(~np.isfinite(ref["poses"]).all(axis=(1, 2)))[0]
    print("   oracle first bad frame", rb[:1], "oracle inliers", ref["inliers"][max(0,first_bad-3):first_bad+1].tolist(), "world", len(ref["world"].xyz), out["world_cnt"][s])
    k = first_bad - 1
    print("   pose diff before", np.abs(out["poses"][s, :k] - ref["poses"][:k]).max())
