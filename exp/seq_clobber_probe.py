"""Round-2 probe of the round-1 'identity array clobbered' observation (profiles/r01_sequences.md, VERDICT weak #3):
runs the bundled dataset as a batch of one and 16 simulated sequences through seq_pipeline_kernel with the library
named by VO_B200_LIB and prints a bit-level checksum of every output, so that the constant-memory build and the
per-thread-array build (-DVO_SEQ_LOCAL_IDENTITY) can be compared, plain and under compute-sanitizer memcheck."""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import replay  # noqa: E402
import simulator  # noqa: E402

vo = importlib.import_module("02-visualodometry_b200")
ctx = vo.Context(0)
ds = dict(np.load(os.path.join(ROOT, "tests", "golden", "dataset.npz")))
F, P = 121, 128
cnt = np.diff(ds["frame_offsets"]).astype(np.int32)[None]
uv = np.zeros((1, F, P, 2), np.float32); desc = np.zeros((1, F, P, 10), np.float32); ids = np.full((1, F, P), -1, np.int32)
for f in range(F):
    fr = replay.frame(ds, f); n = len(fr["uv"])
    uv[0, f, :n], desc[0, f, :n], ids[0, f, :n] = fr["uv"], fr["desc"], fr["id_real"]
out = ctx.seq_batch_run(vo.seq_params(replay.K_REF), cnt, uv, desc, ids)
b = simulator.make_batch(list(range(42, 58)), n_frames=30)
out2 = ctx.seq_batch_run(vo.seq_params(replay.K_REF), b["cnt"], b["uv"], b["desc"], b["id_real"])
h = hashlib.sha256()
for o in (out, out2):
    for k in ("poses", "world_cnt", "world_id", "status", "rounds"):
        h.update(np.ascontiguousarray(o[k]).tobytes())
print("lib", os.path.basename(vo.LIB_PATH), "status", out["status"].tolist(), out2["status"].tolist(), "world", int(out["world_cnt"][0]),
      "pose[1]", out["poses"][0, 1].ravel()[:4].tolist(), "sha256", h.hexdigest()[:16])
