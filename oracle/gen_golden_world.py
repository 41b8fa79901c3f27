#!/usr/bin/env python
"""Ground truth of the bundled simulator for the reference's manual KATs (exec/triangulate_points_test.cpp:33-72,
exec/pose_recovery_test.cpp:29-62): data/world.dat (1000 landmarks: id, xyz, 10-float descriptor) and the
camera-in-robot transform of data/camera.dat.  Run in the BUILD container only (reads /root/reference); writes
tests/golden/world_gt.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main(ref="/root/reference"):
    rows = np.loadtxt(os.path.join(ref, "data", "world.dat"))
    assert rows.shape == (1000, 14)
    cam = open(os.path.join(ref, "data", "camera.dat")).read().split("\n")
    i = cam.index("cam_transform:")
    T = np.array([[float(x) for x in cam[i + 1 + r].split()] for r in range(4)])
    path = os.path.join(ROOT, "tests", "golden", "world_gt.npz")
    np.savez_compressed(path, id=rows[:, 0].astype(np.int32), xyz=rows[:, 1:4].astype(np.float64),
                        desc=rows[:, 4:14].astype(np.float32), cam_in_robot=T)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(*sys.argv[1:])
