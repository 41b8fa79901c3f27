#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (run in the BUILD
container only: needs /root/reference and cv2; neither exists on the GPU box).

  dataset.npz       the bundled 121-frame dataset (reference data/meas-*.dat, parsed
                    as src/my_utilities.cpp:35-112 does) + the reference's own goldens
                    output/{estimated_trajectory,estimated_trajectory_scaled,errors,
                    estimated_world_points}.txt (written by exec/icp_test.cpp:147-210)
  cv2_fixtures.npz  black-box outputs of cv2 (the only OpenCV in this image) for the
                    calls the reference makes at src/cam.cpp:49,61,115,118:
                    findEssentialMat(RANSAC) + recoverPose, triangulatePoints +
                    convertPointsFromHomogeneous, on dataset frame pairs and on
                    seeded synthetic two-view problems.

Usage: python oracle/gen_golden.py [--ref /root/reference]
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

K = np.array([[180, 0, 320], [0, 180, 240], [0, 0, 1]], np.float32)  # src/cam.cpp:11-16


def parse_meas(path):
    gt = odom = None
    rows = []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "gt_pose:":
                gt = [np.float32(t) for t in tok[1:4]]
            elif tok[0] == "odom_pose:":
                odom = [np.float32(t) for t in tok[1:4]]
            elif tok[0] == "point" and len(tok) >= 15:
                rows.append((int(tok[1]), int(tok[2]), [np.float32(t) for t in tok[3:5]],
                             [np.float32(t) for t in tok[5:15]]))
    return gt, odom, rows


def load_dataset(ref):
    offs = [0]
    id_meas, id_real, uv, desc, gts, odoms = [], [], [], [], [], []
    for i in range(121):
        gt, odom, rows = parse_meas(os.path.join(ref, "data", "meas-%05d.dat" % i))
        gts.append(gt)
        odoms.append(odom)
        for r in rows:
            id_meas.append(r[0])
            id_real.append(r[1])
            uv.append(r[2])
            desc.append(r[3])
        offs.append(len(id_meas))
    return dict(frame_offsets=np.array(offs, np.int32), id_meas=np.array(id_meas, np.int32),
                id_real=np.array(id_real, np.int32), uv=np.array(uv, np.float32),
                desc=np.array(desc, np.float32), gt_pose=np.array(gts, np.float32),
                odom_pose=np.array(odoms, np.float32))


def load_outputs(ref):
    o = os.path.join(ref, "output")
    return dict(golden_traj=np.loadtxt(os.path.join(o, "estimated_trajectory.txt")),
                golden_traj_scaled=np.loadtxt(os.path.join(o, "estimated_trajectory_scaled.txt")),
                golden_errors=np.loadtxt(os.path.join(o, "errors.txt")),
                golden_world_points=np.loadtxt(os.path.join(o, "estimated_world_points.txt")))


def frame(ds, i):
    a, b = ds["frame_offsets"][i], ds["frame_offsets"][i + 1]
    return dict(id_meas=ds["id_meas"][a:b], id_real=ds["id_real"][a:b], uv=ds["uv"][a:b], desc=ds["desc"][a:b])


def id_join(f0, f1):
    """matches by ground-truth id (on this dataset identical to descriptor matching)."""
    pos = {int(r): j for j, r in enumerate(f1["id_real"])}
    return np.array([(i, pos[int(r)]) for i, r in enumerate(f0["id_real"]) if int(r) in pos], np.int32)


def cv2_two_view(cv2, x1, x2):
    cv2.setRNGSeed(42)  # src/cam.cpp:40
    E, rmask = cv2.findEssentialMat(x1, x2, K, cv2.RANSAC)
    good, R, t, mask = cv2.recoverPose(E, x1, x2, K)
    return E, rmask.ravel(), R, t.ravel(), mask.ravel(), good


def cv2_triangulate(cv2, T1inv, T2inv, x1, x2):
    """src/cam.cpp:108-118 with T*inv = (camera-in-world pose)^-1 as float32 4x4."""
    P1 = cv2.gemm(K, np.ascontiguousarray(T1inv[:3, :4], np.float32), 1.0, None, 0.0)
    P2 = cv2.gemm(K, np.ascontiguousarray(T2inv[:3, :4], np.float32), 1.0, None, 0.0)
    X4 = cv2.triangulatePoints(P1, P2, np.ascontiguousarray(x1.T), np.ascontiguousarray(x2.T)).T
    X3 = cv2.convertPointsFromHomogeneous(np.ascontiguousarray(X4)).reshape(-1, 3)
    return P1, P2, X4.astype(np.float32), X3.astype(np.float32)


def synth_two_view(seed, n, noise_px):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(-0.15, 0.15, 3)
    cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    R = Rx @ Ry @ Rz
    t = rng.normal(size=3)
    t /= np.linalg.norm(t)
    t *= 0.3
    u = rng.uniform(20, 620, n)
    v = rng.uniform(20, 460, n)
    z = rng.uniform(1.0, 6.0, n)
    Kd = K.astype(np.float64)
    X1 = np.stack([(u - Kd[0, 2]) / Kd[0, 0] * z, (v - Kd[1, 2]) / Kd[1, 1] * z, z], 1)
    X2 = X1 @ R.T + t
    x2 = (X2 @ Kd.T)
    x2 = x2[:, :2] / x2[:, 2:3]
    x1 = np.stack([u, v], 1) + rng.normal(scale=noise_px, size=(n, 2))
    x2 = x2 + rng.normal(scale=noise_px, size=(n, 2))
    keep = (X2[:, 2] > 0.2) & (x2[:, 0] > 0) & (x2[:, 0] < 639) & (x2[:, 1] > 0) & (x2[:, 1] < 479)
    return x1[keep].astype(np.float32), x2[keep].astype(np.float32), R, t / np.linalg.norm(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    import cv2

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    ds = load_dataset(args.ref)
    ds.update(load_outputs(args.ref))
    np.savez_compressed(os.path.join(out_dir, "dataset.npz"), **ds)
    print("dataset.npz:", {k: v.shape for k, v in ds.items()})

    fx = {"cv2_version": np.array(cv2.__version__), "K": K}
    # --- dataset frame pairs: essential + recoverPose, then triangulation with that pose
    pairs = [(0, 1), (1, 2), (10, 11), (37, 38), (60, 61), (99, 100), (119, 120)]
    fx["ds_pairs"] = np.array(pairs, np.int32)
    for n, (a, b) in enumerate(pairs):
        f0, f1 = frame(ds, a), frame(ds, b)
        m = id_join(f0, f1)
        x1, x2 = f0["uv"][m[:, 0]], f1["uv"][m[:, 1]]
        E, rmask, R, t, mask, good = cv2_two_view(cv2, x1, x2)
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = R.astype(np.float32)
        T[:3, 3] = t.astype(np.float32)
        # reference: pose2 (camera-in-world) = [R|t]^-1, T2inv = [R|t]; T1 = identity
        P1, P2, X4, X3 = cv2_triangulate(cv2, np.eye(4, dtype=np.float32), T, x1, x2)
        fx.update({f"ds{n}_x1": x1, f"ds{n}_x2": x2, f"ds{n}_E": E, f"ds{n}_ransac_mask": rmask,
                   f"ds{n}_R": R, f"ds{n}_t": t, f"ds{n}_mask": mask, f"ds{n}_good": np.array(good),
                   f"ds{n}_T2inv": T, f"ds{n}_P1": P1, f"ds{n}_P2": P2, f"ds{n}_X4": X4, f"ds{n}_X3": X3})
        print(f"pair {a}-{b}: {len(m)} matches, good={good}, |t|={np.linalg.norm(t):.6f}")
    # --- synthetic two-view problems (general rotation, optional pixel noise)
    syn = [(1, 200, 0.0), (2, 500, 0.0), (3, 300, 0.05), (4, 1000, 0.2), (5, 64, 0.0)]
    fx["syn_cfg"] = np.array(syn, np.float64)
    for n, (seed, npts, noise) in enumerate(syn):
        x1, x2, Rgt, tgt = synth_two_view(seed, npts, noise)
        E, rmask, R, t, mask, good = cv2_two_view(cv2, x1, x2)
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = R.astype(np.float32)
        T[:3, 3] = t.astype(np.float32)
        P1, P2, X4, X3 = cv2_triangulate(cv2, np.eye(4, dtype=np.float32), T, x1, x2)
        fx.update({f"syn{n}_x1": x1, f"syn{n}_x2": x2, f"syn{n}_E": E, f"syn{n}_ransac_mask": rmask,
                   f"syn{n}_R": R, f"syn{n}_t": t, f"syn{n}_mask": mask, f"syn{n}_good": np.array(good),
                   f"syn{n}_Rgt": Rgt, f"syn{n}_tgt": tgt, f"syn{n}_T2inv": T, f"syn{n}_P1": P1,
                   f"syn{n}_P2": P2, f"syn{n}_X4": X4, f"syn{n}_X3": X3})
        print(f"syn {seed}: n={len(x1)} noise={noise} good={good} "
              f"dR={np.abs(R - Rgt).max():.2e} dt={np.abs(t - tgt).max():.2e}")
    np.savez_compressed(os.path.join(out_dir, "cv2_fixtures.npz"), **fx)
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
