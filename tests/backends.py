"""Backends for tests/replay.py: the CPU oracle, and the CUDA product through its C-ABI."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402


def product():
    """The product package (its directory name is not a Python identifier)."""
    return importlib.import_module("02-visualodometry_b200")


class OracleBackend:
    name = "oracle"

    def __init__(self, essential="ransac", cv2_first_pose=None):
        self.essential = essential
        self.cv2_first_pose = cv2_first_pose

    def match(self, dA, dB, idA=None, idB=None):
        if len(dA) == 0 or len(dB) == 0:
            return np.zeros((0, 2), np.int32), (0, 0)
        return O.match(dA, dB, 0.2, 0.8, idA, idB)

    def essential_recover(self, K, x1, x2):
        if self.cv2_first_pose is not None:  # anchor run: cv2's own (R, t) from the fixtures
            R, t, mask = self.cv2_first_pose
            return R, t, mask
        E, R, t, mask, good = O.essential_recover(K, x1, x2, method=self.essential)
        return R, t, mask

    def triangulate(self, K, T1, T2, x1, x2):
        return O.triangulate(K, T1, T2, x1, x2)

    def pose_inverse(self, T):
        return O.pose_inverse(T)

    def anti_join(self, matched_ids, cand_ids):
        return O.anti_join(matched_ids, cand_ids)

    def picp_init(self, K, rows, cols, pose, world, image, pairs):
        return dict(K=K, rows=rows, cols=cols, pose=np.array(pose, np.float32), world=np.array(world, np.float32),
                    image=np.array(image, np.float32), pairs=np.array(pairs, np.int32))

    def picp_one_round(self, s, thr, damping, keep_outliers):
        s["pose"], ci, co, ni = O.one_round(s["K"], s["rows"], s["cols"], s["pose"], s["world"], s["image"],
                                            s["pairs"], thr, damping, keep_outliers)
        return ci, co, ni

    def picp_pose(self, s):
        return s["pose"]

    def picp_free(self, s):
        pass


class GpuBackend:
    """Every numeric step goes through libvo_b200.so (host buffers in, host buffers out)."""
    name = "gpu"

    def __init__(self, ctx=None, essential="ransac"):
        self.vo = product()
        self.ctx = ctx or self.vo.Context(0)
        self.essential = essential

    def match(self, dA, dB, idA=None, idB=None):
        if len(dA) == 0 or len(dB) == 0:
            return np.zeros((0, 2), np.int32), (0, 0)
        return self.ctx.match(dA, dB, 0.2, 0.8, idA, idB)

    def essential_recover(self, K, x1, x2):
        E, R, t, mask, good = self.ctx.essential_recover(K, x1, x2, method=self.essential)
        return R, t, mask

    def triangulate(self, K, T1, T2, x1, x2):
        return self.ctx.triangulate(K, T1, T2, x1, x2)

    def pose_inverse(self, T):
        return self.vo.pose_inverse(T)

    def anti_join(self, matched_ids, cand_ids):
        return self.ctx.anti_join(matched_ids, cand_ids)

    def picp_init(self, K, rows, cols, pose, world, image, pairs):
        s = self.ctx.picp()
        s.set_camera(K, rows, cols, pose)
        s.set_points(world, image)
        s.set_correspondences(pairs)
        return s

    def picp_one_round(self, s, thr, damping, keep_outliers):
        st = s.one_round(thr, damping, keep_outliers)
        return st.chi_inliers, st.chi_outliers, st.num_inliers

    def picp_pose(self, s):
        return s.get_pose()

    def picp_free(self, s):
        s.close()
