#!/usr/bin/env python
"""More black-box cv2 fixtures for src/cam.cpp:94-140 (triangulatePoints + convertPointsFromHomogeneous) on
two-view problems the bundled dataset does not contain: both cameras away from the origin, scene scales
0.01..100, baselines from 1e-4 of the scene depth to wide, pixel noise up to 2 px, points almost at infinity.
Run in the BUILD container only (needs cv2); writes tests/golden/cv2_triangulate.npz.

Usage: python oracle/gen_golden_tri.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)


def main():
    import cv2
    import synth
    import pyoracle as O
    from gen_golden import K, cv2_triangulate
    rng = np.random.default_rng(20260101)
    out = dict(cv2_version=cv2.__version__, K=K, n_cases=np.int32(24))
    for case in range(24):
        n = int(rng.choice([8, 60, 300]))
        scale = float(rng.choice([0.01, 1.0, 1.0, 100.0]))
        base = rng.normal(0, 1, 3)
        base *= scale * float(rng.choice([1e-4, 0.02, 0.2, 2.0])) / np.linalg.norm(base)
        T1 = synth.euler_pose(np.array([*(rng.normal(0, 1, 3) * scale), *(rng.normal(0, 0.3, 3))]))
        rel = synth.euler_pose(np.array([*base, *(rng.normal(0, 1, 3) * float(rng.choice([0.0, 0.02, 0.3])))]))
        T2 = np.zeros((3, 4))
        T2[:, :3] = T1[:, :3] @ rel[:, :3]
        T2[:, 3] = T1[:, :3] @ rel[:, 3] + T1[:, 3]
        Xc = np.stack([rng.normal(0, 1.5, n), rng.normal(0, 1.0, n), rng.uniform(0.5, 30, n)], 1) * scale
        Xc[rng.random(n) < 0.05, 2] *= 1e5
        X = Xc @ T1[:, :3].T + T1[:, 3]

        def proj(T):
            c = (X - T[:, 3]) @ T[:, :3]
            q = c @ K.astype(np.float64).T
            return q[:, :2] / q[:, 2:3]
        noise = float(rng.choice([0.0, 0.3, 2.0]))
        x1 = (proj(T1) + rng.normal(0, noise, (n, 2))).astype(np.float32)
        x2 = (proj(T2) + rng.normal(0, noise, (n, 2))).astype(np.float32)
        T1f, T2f = T1.astype(np.float32), T2.astype(np.float32)
        T1inv = np.vstack([O.pose_inverse(T1f), [0, 0, 0, 1]]).astype(np.float32)   # Isometry3f::inverse(), float32
        T2inv = np.vstack([O.pose_inverse(T2f), [0, 0, 0, 1]]).astype(np.float32)
        P1, P2, X4, X3 = cv2_triangulate(cv2, T1inv, T2inv, x1, x2)
        out.update({f"c{case}_T1": T1f, f"c{case}_T2": T2f, f"c{case}_x1": x1, f"c{case}_x2": x2, f"c{case}_X3": X3,
                    f"c{case}_X4": X4, f"c{case}_cfg": np.array([scale, np.linalg.norm(base), noise], np.float64)})
    path = os.path.join(ROOT, "tests", "golden", "cv2_triangulate.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
