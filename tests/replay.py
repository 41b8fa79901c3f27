"""Replay of the reference's final pipeline (exec/icp_test.cpp:17-215) over a backend.

The control flow below follows exec/icp_test.cpp line by line; every numeric step
is delegated to `backend`, which is either the CPU oracle (tests/backends.OracleBackend)
or the CUDA product through its C-ABI (tests/backends.GpuBackend).  Used to pin the
oracle against the reference's own goldens (output/*.txt -> tests/golden/dataset.npz)
and to check the CUDA path end to end.
"""
import numpy as np

K_REF = np.array([[180, 0, 320], [0, 180, 240], [0, 0, 1]], np.float32)  # src/cam.cpp:11-16
ROWS, COLS = 480, 640  # exec/icp_test.cpp:30
CAM_TO_IMAGE = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], np.float32)  # src/cam.cpp:18-27
I34 = np.eye(4, dtype=np.float32)[:3].copy()


def frame(ds, i):
    a, b = int(ds["frame_offsets"][i]), int(ds["frame_offsets"][i + 1])
    return dict(id_meas=ds["id_meas"][a:b], id_real=ds["id_real"][a:b], uv=ds["uv"][a:b], desc=ds["desc"][a:b])


class World:
    """std::vector<World_Point> (src/data_point.h:17-31) as parallel arrays."""

    def __init__(self):
        self.xyz = np.zeros((0, 3), np.float32)
        self.desc = np.zeros((0, 10), np.float32)
        self.id_real = np.zeros(0, np.int32)
        self.id_meas = np.zeros(0, np.int32)

    def append(self, xyz, src, idx):  # src/cam.cpp:122-139: descriptor/ids of the FIRST view
        self.xyz = np.concatenate([self.xyz, xyz.astype(np.float32).reshape(-1, 3)])
        self.desc = np.concatenate([self.desc, src["desc"][idx]])
        self.id_real = np.concatenate([self.id_real, src["id_real"][idx]])
        self.id_meas = np.concatenate([self.id_meas, src["id_meas"][idx]])


def picp_frame(backend, pose_wic, world_xyz, image_uv, pairs, thr=3000.0, max_iters=50, conv=1e-5,
               keep_outliers=False):
    """exec/icp_test.cpp:81-111: init, thr 3000, <=50 rounds, stop on 1e-5 relative chi change."""
    solver = backend.picp_init(K_REF, ROWS, COLS, pose_wic, world_xyz, image_uv, pairs)
    prev = np.float32(np.finfo(np.float32).max)
    iters = 0
    n_inl = 0
    for _ in range(max_iters):
        chi_in, chi_out, n_inl = backend.picp_one_round(solver, thr, 1.0, keep_outliers)
        iters += 1
        cur = np.float32(chi_in)
        rel = np.float32(abs(prev - cur)) / prev if prev > 1e-10 else np.float32(0)
        if rel < np.float32(conv):
            break
        prev = cur
    pose = backend.picp_pose(solver)
    backend.picp_free(solver)
    return pose, iters, n_inl


def run_icp_test(ds, backend, n_meas=121, log=None):
    f0, f1 = frame(ds, 0), frame(ds, 1)
    poses = [I34.copy()]
    m01, _ = backend.match(f0["desc"], f1["desc"], f0["id_real"], f1["id_real"])
    R, t, mask = backend.essential_recover(K_REF, f0["uv"][m01[:, 0]], f1["uv"][m01[:, 1]])
    T = np.concatenate([R.astype(np.float32), t.astype(np.float32).reshape(3, 1)], 1)  # cv2eigen casts
    initial_pose = backend.pose_inverse(T)  # src/cam.cpp:78-81 + getPose()
    world = World()
    xyz = backend.triangulate(K_REF, I34, initial_pose, f0["uv"][m01[:, 0]], f1["uv"][m01[:, 1]])
    world.append(xyz, f0, m01[:, 0])
    iters_log, inl_log = [], []
    for i in range(n_meas - 1):
        curr, nxt = frame(ds, i), frame(ds, i + 1)
        iw, _ = backend.match(nxt["desc"], world.desc, nxt["id_real"], world.id_real)
        prev_pose = poses[-1]
        pose_wic, iters, n_inl = picp_frame(backend, backend.pose_inverse(prev_pose), world.xyz, nxt["uv"], iw)
        est = backend.pose_inverse(pose_wic)
        poses.append(est)
        iters_log.append(iters)
        inl_log.append((n_inl, len(iw)))
        im, _ = backend.match(curr["desc"], nxt["desc"], curr["id_real"], nxt["id_real"])
        keep = backend.anti_join(nxt["id_meas"][iw[:, 0]], nxt["id_meas"][im[:, 1]])
        new = im[keep]
        if len(new):
            xyz = backend.triangulate(K_REF, prev_pose, est, curr["uv"][new[:, 0]], nxt["uv"][new[:, 1]])
            world.append(xyz, curr, new[:, 0])
        if log:
            log(i, iters, n_inl, len(iw), len(world.xyz))
    return dict(poses=np.stack(poses), world=world, iters=np.array(iters_log), inliers=np.array(inl_log),
                n_init_matches=len(m01), init_mask=mask)


def run_vo(ds, backend, n_meas=120):
    """The reference's older driver (exec/vo.cpp:55-214): Cam::initOneRound / Cam::oneRound = kernel threshold 1000,
    exactly five rounds (src/cam.cpp:178-224), in the driver's own pose convention (getPose() = world-in-camera is
    stored and handed to triangulatePoints as is)."""
    poses = [I34.copy()]
    world = World()
    n_inl = []
    for i in range(n_meas - 1):
        p1, p2 = frame(ds, i), frame(ds, i + 1)
        if i == 0:
            m, _ = backend.match(p1["desc"], p2["desc"], p1["id_real"], p2["id_real"])
            R, t, _ = backend.essential_recover(K_REF, p1["uv"][m[:, 0]], p2["uv"][m[:, 1]])
            T = np.concatenate([R.astype(np.float32), t.astype(np.float32).reshape(3, 1)], 1)
            est = backend.pose_inverse(T)
            world.append(backend.triangulate(K_REF, I34, est, p1["uv"][m[:, 0]], p2["uv"][m[:, 1]]), p1, m[:, 0])
            poses.append(est)
            continue
        iw, _ = backend.match(p2["desc"], world.desc, p2["id_real"], world.id_real)
        prev = poses[-1]
        solver = backend.picp_init(K_REF, ROWS, COLS, prev, world.xyz, p2["uv"], iw)
        inl = 0
        for _ in range(5):
            _, _, inl = backend.picp_one_round(solver, 1000.0, 1.0, False)
        est = backend.picp_pose(solver)
        backend.picp_free(solver)
        poses.append(est)
        n_inl.append((inl, len(iw)))
        im, _ = backend.match(p1["desc"], p2["desc"], p1["id_real"], p2["id_real"])
        keep = backend.anti_join(p2["id_meas"][iw[:, 0]], p2["id_meas"][im[:, 1]])
        new = im[keep]
        if len(new):
            world.append(backend.triangulate(K_REF, prev, est, p1["uv"][new[:, 0]], p2["uv"][new[:, 1]]), p1, new[:, 0])
    return dict(poses=np.stack(poses), world=world, inliers=np.array(n_inl))


def augment_pose(p):  # src/my_utilities.cpp:245-260
    T = np.eye(4, dtype=np.float64)[:3]
    c, s = np.cos(np.float32(p[2])), np.sin(np.float32(p[2]))
    T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    T[0, 3], T[1, 3] = p[0], p[1]
    return T


def umeyama_scale(P, Q):
    """Eigen::umeyama(P, Q, true) scale factor (src/my_utilities.cpp:459-478); P,Q are N x 3."""
    P = np.asarray(P, np.float64)
    Q = np.asarray(Q, np.float64)
    mp, mq = P.mean(0), Q.mean(0)
    Pd, Qd = P - mp, Q - mq
    src_var = (Pd ** 2).sum() / len(P)
    sigma = Qd.T @ Pd / len(P)
    U, d, Vt = np.linalg.svd(sigma)
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1
    return float((d * S).sum() / src_var)


def evaluate(ds, res):
    """exec/icp_test.cpp:138-210: re-frame by cameraToImage, umeyama scale, the four output tables."""
    poses = res["poses"].astype(np.float64)
    n = len(poses)
    C = CAM_TO_IMAGE.astype(np.float64)
    Rs = np.einsum("ij,njk->nik", C, poses[:, :, :3])
    ts = np.einsum("ij,nj->ni", C, poses[:, :, 3])
    gt = np.stack([augment_pose(ds["gt_pose"][j]) for j in range(n)])
    scale = umeyama_scale(ts, gt[:, :, 3])
    angle = np.arctan2(Rs[:, 1, 0], Rs[:, 0, 0]).astype(np.float32) + np.float32(np.pi / 2.0)
    angle_gt = np.arctan2(gt[:, 1, 0], gt[:, 0, 0])
    idx = np.arange(n)
    traj = np.stack([idx, ts[:, 0], ts[:, 1], angle], 1)
    ts_s = ts * scale
    traj_s = np.stack([idx, ts_s[:, 0], ts_s[:, 1], angle], 1)
    err = np.stack([idx, np.linalg.norm(ts_s - gt[:, :, 3], axis=1), np.abs(angle - angle_gt)], 1)
    w = res["world"]
    rows = []
    seen = set()
    order = np.argsort(w.id_real, kind="stable")
    for k in order:  # first world point per id_real, ids ascending (icp_test.cpp:199-210)
        rid = int(w.id_real[k])
        if rid in seen or not (0 <= rid < 1000):
            continue
        seen.add(rid)
        p = C @ w.xyz[k].astype(np.float64) * scale
        rows.append([rid, p[0], p[1], p[2]])
    return dict(scale=scale, traj=traj, traj_scaled=traj_s, errors=err, world_points=np.array(rows))


def _robot_T(p):
    x, y, th = p
    T = np.eye(4)
    T[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
    T[0, 3], T[1, 3] = x, y
    return T


def triangulation_kat(be, dataset, world_gt):
    """exec/triangulate_points_test.cpp:33-72 as a known-answer test: frames 0/1 -> match -> essential + recoverPose ->
    triangulate, against data/world.dat[id_real] moved into camera 0 (ground-truth robot pose x camera-in-robot of
    data/camera.dat).  The reconstruction is in units of the baseline (|t| = 1): the least-squares scale must be the
    true baseline 0.2004 and every landmark within 0.5 % of its distance (measured: median 3.6e-4, max 2.3e-3; the
    pixels carry 1e-2 px print-precision noise)."""
    f0, f1 = frame(dataset, 0), frame(dataset, 1)
    m, _ = be.match(f0["desc"], f1["desc"], f0["id_real"], f1["id_real"])
    x1, x2 = f0["uv"][m[:, 0]], f1["uv"][m[:, 1]]
    R, t, mask = be.essential_recover(K_REF, x1, x2)
    assert len(m) == 115 and int((np.asarray(mask) != 0).sum()) == 115
    T = np.concatenate([np.asarray(R, np.float32), np.asarray(t, np.float32).reshape(3, 1)], 1)
    X = be.triangulate(K_REF, I34, be.pose_inverse(T), x1, x2).astype(np.float64)
    ids = f0["id_real"][m[:, 0]]
    C0 = _robot_T(dataset["gt_pose"][0]) @ world_gt["cam_in_robot"]
    C1 = _robot_T(dataset["gt_pose"][1]) @ world_gt["cam_in_robot"]
    Xw = np.concatenate([world_gt["xyz"][ids], np.ones((len(ids), 1))], 1)
    Xc = (np.linalg.inv(C0) @ Xw.T).T[:, :3]
    base = np.linalg.norm(C1[:3, 3] - C0[:3, 3])
    scale = (X * Xc).sum() / (X * X).sum()
    assert abs(scale - base) <= 1e-3 * base, (scale, base)
    err = np.linalg.norm(X * base - Xc, axis=1) / np.linalg.norm(Xc, axis=1)
    assert err.max() <= 5e-3 and np.median(err) <= 1e-3, (err.max(), np.median(err))
