"""Synthetic sequences with the statistics of the bundled dataset:
1000 landmarks U(-10,10)^2 x U(0,2) with U(-1,1)^10 descriptors, a planar robot moving 0.2 units per
frame, the camera of data/camera.dat (K = [180 0 320; 0 180 240; 0 0 1], camera-in-robot
R = [0 0 1; -1 0 0; 0 -1 0], t = (0.2, 0, 0)), visibility 0 < z < 5 inside 640x480. numpy only."""
import numpy as np

K = np.array([[180, 0, 320], [0, 180, 240], [0, 0, 1]], np.float64)
R_RC = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], np.float64)  # camera axes in the robot frame
T_RC = np.array([0.2, 0.0, 0.0])
MAX_PTS = 128


def make_sequence(seed, n_frames=121, n_landmarks=1000, max_pts=MAX_PTS):
    """returns dict(cnt[F], uv[F,P,2], desc[F,P,10], id_real[F,P], gt_pose[F,3]) (float32 / int32)"""
    rng = np.random.Generator(np.random.Philox(seed))
    lm = np.stack([rng.uniform(-10, 10, n_landmarks), rng.uniform(-10, 10, n_landmarks), rng.uniform(0, 2, n_landmarks)], 1)
    ldesc = rng.uniform(-1, 1, (n_landmarks, 10)).astype(np.float32)
    # planar path: constant 0.2 forward step, slowly varying turn rate (first step straight, like the dataset)
    om = np.cumsum(rng.normal(0, 0.01, n_frames)) * 0.3
    om[:2] = 0
    x = y = th = 0.0
    gt = np.zeros((n_frames, 3))
    for f in range(n_frames):
        gt[f] = (x, y, th)
        th += np.clip(om[f], -0.08, 0.08)
        x += 0.2 * np.cos(th)
        y += 0.2 * np.sin(th)
    cnt = np.zeros(n_frames, np.int32)
    uv = np.zeros((n_frames, max_pts, 2), np.float32)
    desc = np.zeros((n_frames, max_pts, 10), np.float32)
    ids = np.full((n_frames, max_pts), -1, np.int32)
    for f in range(n_frames):
        c, s = np.cos(gt[f, 2]), np.sin(gt[f, 2])
        Rwr = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        pr = (lm - np.array([gt[f, 0], gt[f, 1], 0.0])) @ Rwr  # robot frame
        pc = (pr - T_RC) @ R_RC                                # camera frame
        z = pc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = K[0, 0] * pc[:, 0] / z + K[0, 2]
            v = K[1, 1] * pc[:, 1] / z + K[1, 2]
        vis = np.nonzero((z > 0) & (z < 5) & (u >= 0) & (u < 640) & (v >= 0) & (v < 480))[0][:max_pts]
        n = len(vis)
        cnt[f] = n
        uv[f, :n, 0], uv[f, :n, 1] = u[vis], v[vis]
        desc[f, :n] = ldesc[vis]
        ids[f, :n] = vis
    return dict(cnt=cnt, uv=uv, desc=desc, id_real=ids, gt_pose=gt.astype(np.float32))


def make_batch(seeds, n_frames=121, max_pts=MAX_PTS):
    seqs = [make_sequence(s, n_frames, max_pts=max_pts) for s in seeds]
    return {k: np.stack([q[k] for q in seqs]) for k in seqs[0]}


def as_dataset(batch, s):
    """one sequence of a batch in the layout of tests/golden/dataset.npz (for tests/replay.py)"""
    cnt = batch["cnt"][s]
    offs = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    F = len(cnt)
    take = lambda a: np.concatenate([a[s, f, :cnt[f]] for f in range(F)])
    return dict(frame_offsets=offs, uv=take(batch["uv"]), desc=take(batch["desc"]), id_real=take(batch["id_real"]),
                id_meas=np.concatenate([np.arange(c, dtype=np.int32) for c in cnt]), gt_pose=batch["gt_pose"][s])
