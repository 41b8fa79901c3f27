#!/usr/bin/env python
"""More black-box cv2 fixtures for src/cam.cpp:49,61 (findEssentialMat(RANSAC) + recoverPose) on synthetic two-view
problems: rotations up to 0.5 rad, baselines in every direction (sideways, forward, backward), 0..1 px noise, 10 %
gross outliers, 20..400 points.  Run in the BUILD container only (needs cv2); writes tests/golden/cv2_recoverpose.npz.
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
    from gen_golden import K, cv2_two_view
    rng = np.random.default_rng(20260102)
    out = dict(cv2_version=cv2.__version__, K=K, n_cases=np.int32(20))
    for case in range(20):
        n = int(rng.choice([20, 100, 400]))
        base = rng.normal(0, 1, 3)
        if case % 4 == 1:
            base = np.array([0.0, 0.0, 1.0])      # forward motion (the dataset's case)
        if case % 4 == 2:
            base = np.array([0.0, 0.0, -1.0])     # backward
        base *= float(rng.choice([0.1, 0.5, 2.0])) / np.linalg.norm(base)
        rel = synth.euler_pose(np.array([*base, *(rng.normal(0, 1, 3) * float(rng.choice([0.0, 0.05, 0.25])))]))
        Xc = np.stack([rng.normal(0, 2.0, n), rng.normal(0, 1.5, n), rng.uniform(2, 20, n)], 1)

        def proj(T):
            c = (Xc - T[:, 3]) @ T[:, :3]
            q = c @ K.astype(np.float64).T
            return q[:, :2] / q[:, 2:3]
        I = np.eye(4)[:3]
        noise = float(rng.choice([0.0, 0.2, 1.0]))
        x1 = (proj(I) + rng.normal(0, noise, (n, 2))).astype(np.float32)
        x2 = (proj(rel) + rng.normal(0, noise, (n, 2))).astype(np.float32)
        bad = rng.random(n) < 0.1
        x2[bad] = rng.uniform(0, 480, (int(bad.sum()), 2)).astype(np.float32)
        E, rmask, R, t, mask, good = cv2_two_view(cv2, x1, x2)
        if E is None or E.shape != (3, 3):
            E = np.zeros((3, 3)); R = np.zeros((3, 3)); t = np.zeros(3); mask = np.zeros(n, np.uint8); good = 0
        out.update({f"c{case}_x1": x1, f"c{case}_x2": x2, f"c{case}_E": np.asarray(E, np.float64), f"c{case}_R": np.asarray(R, np.float64),
                    f"c{case}_t": np.asarray(t, np.float64).ravel(), f"c{case}_mask": np.asarray(mask, np.uint8).ravel(), f"c{case}_good": np.int32(good),
                    f"c{case}_rel": rel})
    path = os.path.join(ROOT, "tests", "golden", "cv2_recoverpose.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
