"""Batched independent sequences (BASELINE config 5): every sequence solved by one CTA on the device must
follow the oracle's replay of exec/icp_test.cpp on the same measurements."""
import numpy as np
import pytest

import backends
import replay
import simulator
from backends import product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    vo = product()
    c = vo.Context(0)
    yield c
    c.close()


def _compare(out, s, ref, n_frames):
    wc = int(out["world_cnt"][s])
    assert out["status"][s] == 0
    # the map: same landmarks in the same order (matching and the anti-join are exact)
    assert wc == len(ref["world"].xyz)
    assert np.array_equal(out["world_id"][s, :wc], ref["world"].id_real)
    assert np.array_equal(out["inliers"][s, 1:n_frames, 1], ref["inliers"][:, 1])
    # poses: the 1e-5 relative stop sits at float noise, so round counts may differ; poses agree
    dp = np.abs(out["poses"][s, :n_frames] - ref["poses"]).max()
    assert dp < 5e-3, dp
    # landmarks: triangulation over a 1-unit baseline amplifies the ~1e-3 pose differences by depth/baseline,
    # so the bound is relative to each point's distance (2 %)
    d = np.linalg.norm(out["world_xyz"][s, :wc] - ref["world"].xyz, axis=1)
    assert (d <= 2e-2 * np.maximum(np.linalg.norm(ref["world"].xyz, axis=1), 1.0)).all(), d.max()


def test_bundled_dataset_as_a_batch_of_one(ctx, dataset):
    """the 121-frame dataset through the batched kernel == the per-call GPU path == the oracle replay"""
    vo = product()
    F, P = 121, 128
    cnt = np.diff(dataset["frame_offsets"]).astype(np.int32)[None]
    uv = np.zeros((1, F, P, 2), np.float32)
    desc = np.zeros((1, F, P, 10), np.float32)
    ids = np.full((1, F, P), -1, np.int32)
    for f in range(F):
        fr = replay.frame(dataset, f)
        n = len(fr["uv"])
        uv[0, f, :n], desc[0, f, :n], ids[0, f, :n] = fr["uv"], fr["desc"], fr["id_real"]
    out = ctx.seq_batch_run(vo.seq_params(replay.K_REF), cnt, uv, desc, ids)
    ref = replay.run_icp_test(dataset, backends.OracleBackend(essential="8pt"))
    _compare(out, 0, ref, F)
    assert out["world_cnt"][0] == 490
    ev = replay.evaluate(dataset, dict(poses=out["poses"][0], world=ref["world"]))
    dxy = np.linalg.norm(ev["traj"][:, 1:3] - dataset["golden_traj"][:, 1:3], axis=1).max()
    assert dxy <= 0.01 * 41.4


def test_synthetic_batch_matches_oracle_replay(ctx):
    vo = product()
    seeds = list(range(42, 42 + 24))
    F = 40
    batch = simulator.make_batch(seeds, n_frames=F)
    out = ctx.seq_batch_run(vo.seq_params(replay.K_REF), batch["cnt"], batch["uv"], batch["desc"], batch["id_real"])
    assert (out["status"] == 0).all()
    for s in (0, 5, 11, 23):
        ds = simulator.as_dataset(batch, s)
        ref = replay.run_icp_test(ds, backends.OracleBackend(essential="8pt"), n_meas=F)
        _compare(out, s, ref, F)
    # independent sequences: a sequence's result does not depend on its neighbours in the batch
    solo = ctx.seq_batch_run(vo.seq_params(replay.K_REF), batch["cnt"][5:6], batch["uv"][5:6], batch["desc"][5:6],
                             batch["id_real"][5:6])
    assert np.array_equal(solo["poses"][0], out["poses"][5]) and solo["world_cnt"][0] == out["world_cnt"][5]


def test_degenerate_sequences(ctx):
    vo = product()
    batch = simulator.make_batch([1, 2], n_frames=6)
    batch["cnt"][1, 0] = 3  # fewer than 8 initial matches: flagged, nothing computed
    out = ctx.seq_batch_run(vo.seq_params(replay.K_REF), batch["cnt"], batch["uv"], batch["desc"], batch["id_real"],
                            world_cap=1024)
    assert out["status"][0] == 0 and out["status"][1] == 2 and out["world_cnt"][1] == 0
    tiny = ctx.seq_batch_run(vo.seq_params(replay.K_REF), batch["cnt"][:1], batch["uv"][:1], batch["desc"][:1],
                             batch["id_real"][:1], world_cap=64)
    assert tiny["status"][0] in (0, 1) and tiny["world_cnt"][0] <= 64
