"""Writes the golden dataset (tests/golden/dataset.npz) back into the reference's on-disk format
(data/meas-NNNNN.dat) so that the native C++ pipeline can be replayed on the GPU
box, where /root/reference does not exist. %.9g round-trips every float32 exactly."""
import os

import numpy as np


def write_meas_files(ds, out_dir, n_meas=121):
    os.makedirs(out_dir, exist_ok=True)
    for i in range(n_meas):
        a, b = int(ds["frame_offsets"][i]), int(ds["frame_offsets"][i + 1])
        with open(os.path.join(out_dir, "meas-%05d.dat" % i), "w") as f:
            f.write("seq: %d\n" % i)
            f.write("gt_pose: %.9g %.9g %.9g\n" % tuple(ds["gt_pose"][i]))
            f.write("odom_pose: %.9g %.9g %.9g\n" % tuple(ds["odom_pose"][i]))
            for k in range(a, b):
                desc = " ".join("%.9g" % x for x in ds["desc"][k])
                f.write("point %d %d %.9g %.9g %s\n" % (ds["id_meas"][k], ds["id_real"][k], ds["uv"][k, 0],
                                                        ds["uv"][k, 1], desc))
    return os.path.join(out_dir, "meas-")


def read_outputs(out_dir):
    return dict(traj=np.loadtxt(os.path.join(out_dir, "estimated_trajectory.txt")),
                traj_scaled=np.loadtxt(os.path.join(out_dir, "estimated_trajectory_scaled.txt")),
                errors=np.loadtxt(os.path.join(out_dir, "errors.txt")),
                world_points=np.loadtxt(os.path.join(out_dir, "estimated_world_points.txt")))


def write_world_file(world_gt, path):
    """data/world.dat of the reference: `id x y z d0..d9` per landmark (src/my_utilities.cpp:137-182)"""
    with open(path, "w") as f:
        for i in range(len(world_gt["id"])):
            desc = " ".join("%.9g" % x for x in world_gt["desc"][i])
            f.write("%d %.9g %.9g %.9g %s\n" % (world_gt["id"][i], *world_gt["xyz"][i], desc))
