"""Multi-GPU tests (need >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multigpu.py -m gpu`;
skipped on a single-GPU box). One process per GPU; torch.distributed only moves the NCCL id and the IPC
handles. Checks: sharded PICP through (a) ncclAllReduce and (b) the fused peer-memory exchange equals the
unsharded single-GPU result (masks/counts exactly, pose to 1e-6) and is identical on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from backends import product

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    vo = product()
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    ctx = vo.Context(rank)
    uid = torch.from_numpy(vo.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, uid.cpu().numpy())
    fr = synth.picp_frame(n=400003, seed=9, permute=True)
    lo, hi = vo.shard_range(len(fr["pairs"]), world, rank)
    res = {}
    for mode in ("nccl", "peer", "peer_resident"):
        if mode == "peer":
            mine = torch.from_numpy(ctx.peer_export()).to(dev)
            allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.peer_attach(world, rank, torch.stack(allh).cpu().numpy())
            assert ctx.peer_active
        s = ctx.picp()
        # "nccl" and "peer": one launch per round (ncclAllReduce / fused mailbox exchange in the last CTA);
        # "peer_resident": all rounds in ONE persistent launch per GPU, the shard resident in shared memory
        s.set_mode(vo.MODE_RESIDENT if mode == "peer_resident" else vo.MODE_STREAM)
        s.set_camera(fr["K"], 480, 640, fr["pose0"])
        s.set_points(fr["world"], fr["image"])
        s.set_correspondences(fr["pairs"][lo:hi])
        lin = s.linearize(3000.0, False)
        dist.barrier()
        s.enqueue_rounds(3000.0, 1.0, False, 6)
        st = s.fetch_stats(6)
        res[mode + "_H"] = lin["H"]
        res[mode + "_n"] = np.array([x.num_inliers for x in st])
        res[mode + "_chi"] = np.array([x.chi_inliers for x in st])
        res[mode + "_pose"] = s.get_pose()
        s.close()
        dist.barrier()
    # sharded matching with its exchange (MAX all-reduce of the per-row results over NCCL): identical on every rank
    A, B = synth.descriptors(30000, 20000, seed=3, noise=0.02)
    dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    midx = torch.empty(30000, dtype=torch.int32, device=dev)
    mp_ = torch.empty((30000, 2), dtype=torch.int32, device=dev)
    nm = ctx.match_sharded_dev(dA.data_ptr(), 30000, dB.data_ptr(), 20000, 10, rank, world, midx.data_ptr(), mp_.data_ptr(), 30000)
    res["match_pairs"] = mp_[:nm].cpu().numpy()
    res["match_idx"] = midx.cpu().numpy()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    ctx.peer_detach()
    ctx.comm_destroy()
    dist.destroy_process_group()
    ctx.close()


def test_sharded_picp_nccl_and_fused_peer_exchange(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    vo = product()
    ctx = vo.Context(0)
    fr = synth.picp_frame(n=400003, seed=9, permute=True)
    s = ctx.picp()
    s.set_mode(vo.MODE_STREAM)
    s.set_camera(fr["K"], 480, 640, fr["pose0"])
    s.set_points(fr["world"], fr["image"])
    s.set_correspondences(fr["pairs"])
    lin = s.linearize(3000.0, False)
    s.enqueue_rounds(3000.0, 1.0, False, 6)
    st = s.fetch_stats(6)
    single_n = np.array([x.num_inliers for x in st])
    for mode in ("nccl", "peer", "peer_resident"):
        # every rank ends with the identical state, no broadcast
        assert np.array_equal(r0[mode + "_pose"], r1[mode + "_pose"])
        assert np.array_equal(r0[mode + "_n"], r1[mode + "_n"]) and np.array_equal(r0[mode + "_chi"], r1[mode + "_chi"])
        assert np.array_equal(r0[mode + "_H"], r1[mode + "_H"])
        # sharded == unsharded
        assert r0[mode + "_n"][0] == single_n[0]
        assert np.abs(r0[mode + "_n"] - single_n).max() <= 4
        assert np.abs(r0[mode + "_H"] - lin["H"]).max() <= 1e-5 * np.abs(lin["H"]).max()
        assert np.abs(r0[mode + "_pose"] - s.get_pose()).max() <= 1e-6
    # sharded matching: both ranks hold the complete, identical match list = the unsharded call
    A, B = synth.descriptors(30000, 20000, seed=3, noise=0.02)
    ref_pairs, _ = ctx.match(A, B)
    assert np.array_equal(r0["match_pairs"], r1["match_pairs"]) and np.array_equal(r0["match_idx"], r1["match_idx"])
    assert np.array_equal(r0["match_pairs"], ref_pairs)
    # the two exchange mechanisms sum the same two numbers: identical bits
    assert np.array_equal(r0["nccl_pose"], r0["peer_pose"])
    s.close()
