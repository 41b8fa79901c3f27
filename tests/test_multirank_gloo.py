"""world_size-2 `gloo` tests of the multi-rank paths on CPU: the partitioning, the 32-term
all-reduce layout and the id-distribution plumbing are exercised with the ORACLE doing each rank's
arithmetic, so what is checked is the multi-rank algorithm itself: sharded == unsharded, and every rank
ends with the identical pose without a broadcast."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from backends import product


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from oracle import pyoracle as O
    vo = product()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # --- plumbing: a 128-byte communicator id made by rank 0 reaches every rank unchanged
    uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    assert uid.tolist() == list(range(128))
    # --- PICP: contiguous correspondence shards, one all-reduce of 32 doubles per round
    fr = synth.picp_frame(n=6001, seed=5, permute=True)
    lo, hi = vo.shard_range(len(fr["pairs"]), world, rank)
    pose = fr["pose0"].copy()
    poses = []
    for _ in range(4):
        part = O.linearize(fr["K"], 480, 640, pose, fr["world"], fr["image"], fr["pairs"][lo:hi], 3000.0, False, accum="f64",
                           want_status=False)
        n_out = 0
        t = torch.from_numpy(vo.pack_terms(part["H"], part["b"], part["chi_in"], part["chi_out"], part["n_inliers"], n_out))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = vo.unpack_terms(t.numpy())
        H = tot["H"].astype(np.float32) + np.eye(6, dtype=np.float32)  # damping 1
        dx = O.ldlt_solve6(H, (-tot["b"]).astype(np.float32))
        pose = O.pose_update(dx, pose)
        poses.append((pose.copy(), tot["n_inliers"], tot["chi_in"]))
    # --- matching: row blocks, no collective
    A, B = synth.descriptors(501, 700, seed=8)
    rlo, rhi = vo.shard_range(len(A), world, rank)
    mine, _ = O.match(A, B, row_begin=rlo, row_end=rhi)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pose=np.stack([p for p, _, _ in poses]),
             n_in=np.array([n for _, n, _ in poses]), chi=np.array([c for _, _, c in poses]), matches=mine)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_picp_and_matching(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # every rank solved the same system: identical poses, no broadcast needed
    assert np.array_equal(r0["pose"], r1["pose"]) and np.array_equal(r0["n_in"], r1["n_in"])
    # sharded == unsharded (float64 accumulation: the only difference is the summation split)
    fr = synth.picp_frame(n=6001, seed=5, permute=True)
    pose = fr["pose0"].copy()
    for k in range(4):
        whole = oracle.linearize(fr["K"], 480, 640, pose, fr["world"], fr["image"], fr["pairs"], 3000.0, False, accum="f64",
                                 want_status=False)
        assert whole["n_inliers"] == r0["n_in"][k]
        assert abs(whole["chi_in"] - r0["chi"][k]) <= 1e-9 * max(whole["chi_in"], 1.0)
        H = whole["H"].astype(np.float32) + np.eye(6, dtype=np.float32)
        pose = oracle.pose_update(oracle.ldlt_solve6(H, (-whole["b"]).astype(np.float32)), pose)
        assert np.abs(pose - r0["pose"][k]).max() <= 1e-6
    A, B = synth.descriptors(501, 700, seed=8)
    full, _ = oracle.match(A, B)
    assert np.array_equal(np.concatenate([r0["matches"], r1["matches"]]), full)


def test_shard_bounds_cover_exactly_once():
    vo = product()
    for n in (0, 1, 7, 8, 1000, 10485760):
        for w in (1, 2, 3, 4, 8):
            b = vo.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_terms_roundtrip():
    vo = product()
    rng = np.random.default_rng(0)
    J = rng.normal(size=(10, 6))
    H = J.T @ J
    t = vo.pack_terms(H, np.arange(6.0), 3.5, 7.25, 11, 4)
    u = vo.unpack_terms(t)
    assert np.allclose(u["H"], H) and np.array_equal(u["b"], np.arange(6.0))
    assert (u["chi_in"], u["chi_out"], u["n_inliers"], u["n_outliers"]) == (3.5, 7.25, 11, 4)
