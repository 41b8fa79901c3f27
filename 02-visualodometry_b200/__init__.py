"""vo-b200: ctypes harness over libvo_b200.so (the C-ABI in include/vo_b200.h).

The product is the shared library (hand-written sm_100a CUDA behind a C-ABI) and its C++
host mirror in host/.  This module only exists so tests/ and bench.py can drive the C-ABI from
Python; it adds no arithmetic and has no fallback: importing fails loudly when the library has
not been built, and every call raises when the library reports an error.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VO_B200_LIB", os.path.join(_HERE, "libvo_b200.so"))  # override: kernel A/B experiments

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C 02-visualodometry_b200` (there is no CPU fallback)")

_L = C.CDLL(LIB_PATH)

_vp = C.c_void_p
_i64 = C.c_int64
_f = C.c_float


class SeqParams(C.Structure):
    _fields_ = [("K", C.c_float * 9), ("rows", C.c_int32), ("cols", C.c_int32), ("dist_thr", C.c_float),
                ("ratio_thr", C.c_float), ("kernel_threshold", C.c_float), ("damping", C.c_float),
                ("keep_outliers", C.c_int32), ("max_rounds", C.c_int32), ("rel_tol", C.c_float)]


def seq_params(K, rows=480, cols=640, dist_thr=0.2, ratio_thr=0.8, kernel_threshold=3000.0, damping=1.0,
               keep_outliers=False, max_rounds=50, rel_tol=1e-5):
    """the constants of exec/icp_test.cpp / src/my_utilities.h as a vo_seq_params"""
    p = SeqParams()
    for i, v in enumerate(np.asarray(K, np.float32).reshape(9)):
        p.K[i] = float(v)
    p.rows, p.cols, p.dist_thr, p.ratio_thr = rows, cols, dist_thr, ratio_thr
    p.kernel_threshold, p.damping, p.keep_outliers = kernel_threshold, damping, int(keep_outliers)
    p.max_rounds, p.rel_tol = max_rounds, rel_tol
    return p


class Stats(C.Structure):
    _fields_ = [("chi_inliers", C.c_float), ("chi_outliers", C.c_float), ("num_inliers", C.c_int32),
                ("num_outliers", C.c_int32)]

    def __repr__(self):
        return (f"Stats(chi_in={self.chi_inliers:.6g}, chi_out={self.chi_outliers:.6g}, "
                f"inliers={self.num_inliers}, outliers={self.num_outliers})")


def _sig(name, restype, *argtypes):
    fn = getattr(_L, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_sig("vo_status_str", C.c_char_p, C.c_int)
_sig("vo_version", C.c_int)
_sig("vo_device_count", C.c_int, C.POINTER(C.c_int))
_sig("vo_ctx_create", C.c_int, C.c_int, _vp, C.POINTER(_vp))
_sig("vo_ctx_destroy", C.c_int, _vp)
_sig("vo_ctx_sync", C.c_int, _vp)
_sig("vo_last_error", C.c_char_p, _vp)
_sig("vo_ctx_kernel_launches", _i64, _vp)
_sig("vo_ctx_stream", _vp, _vp)
_sig("vo_comm_unique_id", C.c_int, _vp)
_sig("vo_ctx_comm_init", C.c_int, _vp, C.c_int, C.c_int, _vp)
_sig("vo_ctx_comm_destroy", C.c_int, _vp)
_sig("vo_ctx_comm_size", C.c_int, _vp)
_sig("vo_ctx_peer_export", C.c_int, _vp, _vp)
_sig("vo_ctx_peer_attach", C.c_int, _vp, C.c_int, C.c_int, _vp)
_sig("vo_ctx_peer_detach", C.c_int, _vp)
_sig("vo_ctx_peer_active", C.c_int, _vp)
_sig("vo_pose_inverse", None, _vp, _vp)
_sig("vo_pose_mul", None, _vp, _vp, _vp)
_sig("vo_project_points", C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _i64, C.c_int, _vp, C.POINTER(_i64),
     C.POINTER(_i64))
_sig("vo_picp_create", C.c_int, _vp, C.POINTER(_vp))
_sig("vo_picp_destroy", C.c_int, _vp)
_sig("vo_picp_set_camera", C.c_int, _vp, _vp, C.c_int, C.c_int, _vp)
_sig("vo_picp_set_pose", C.c_int, _vp, _vp)
_sig("vo_picp_get_pose", C.c_int, _vp, _vp)
_sig("vo_picp_set_points", C.c_int, _vp, _vp, _i64, _vp, _i64)
_sig("vo_picp_set_points_dev", C.c_int, _vp, _vp, _i64, _vp, _i64)
_sig("vo_picp_set_correspondences", C.c_int, _vp, _vp, _i64)
_sig("vo_picp_set_correspondences_dev", C.c_int, _vp, _vp, _i64)
_sig("vo_picp_set_mode", C.c_int, _vp, C.c_int)
_sig("vo_picp_pack", C.c_int, _vp)
_sig("vo_picp_resident_capacity", C.c_int, _vp, C.POINTER(_i64))
_sig("vo_picp_linearize", C.c_int, _vp, _f, C.c_int, _vp, _vp, C.POINTER(Stats), _vp)
_sig("vo_picp_one_round", C.c_int, _vp, _f, _f, C.c_int, C.POINTER(Stats))
_sig("vo_picp_enqueue_rounds", C.c_int, _vp, _f, _f, C.c_int, C.c_int)
_sig("vo_picp_fetch_stats", C.c_int, _vp, _vp, C.c_int)
_sig("vo_picp_solve", C.c_int, _vp, _f, _f, C.c_int, C.c_int, _f, C.POINTER(C.c_int), C.POINTER(Stats))
_sig("vo_match", C.c_int, _vp, _vp, _i64, _vp, _i64, C.c_int, _f, _f, _vp, _vp, _i64, _i64, _vp, _i64,
     C.POINTER(_i64), _vp)
_sig("vo_match_dev", C.c_int, _vp, _vp, _i64, _vp, _i64, C.c_int, _f, _f, _vp, _vp, _i64, _i64, _vp, _i64,
     C.POINTER(_i64), _vp, _vp, _vp, _vp)
_sig("vo_match_sharded_dev", C.c_int, _vp, _vp, _i64, _vp, _i64, C.c_int, _f, _f, C.c_int, C.c_int, _vp, _vp, _i64,
     C.POINTER(_i64))
_sig("vo_match_compact_dev", C.c_int, _vp, _vp, _i64, _vp, _i64, C.POINTER(_i64))
_sig("vo_match_set_path", C.c_int, _vp, C.c_int)
_sig("vo_selftest_reciprocal", C.c_int, _vp, _vp)
_sig("vo_triangulate", C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp)
_sig("vo_triangulate_dev", C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp)
_sig("vo_essential_recover", C.c_int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(C.c_int))
_sig("vo_essential_recover_ex", C.c_int, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_double, C.c_double, C.c_int, _vp, _vp, _vp,
     _vp, C.POINTER(C.c_int), _vp, C.POINTER(C.c_int), C.POINTER(C.c_int))
_sig("vo_anti_join", C.c_int, _vp, _vp, _i64, _vp, _i64, _vp, C.POINTER(_i64))
_sig("vo_seq_batch_run", C.c_int, _vp, C.POINTER(SeqParams), C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp,
     _vp, _vp, _vp, _vp, _vp, _vp, _vp)
_sig("vo_seq_batch_run_dev", C.c_int, _vp, C.POINTER(SeqParams), C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp,
     _vp, _vp, _vp, _vp, _vp, _vp, _vp)

from .sharding import N_TERMS, pack_terms, shard_bounds, shard_range, unpack_terms  # noqa: E402,F401

MAX_ROUNDS = 64
MODE_AUTO, MODE_STREAM, MODE_RESIDENT, MODE_STREAM_PERSISTENT = 0, 1, 2, 3
STATUS_SKIPPED, STATUS_INLIER, STATUS_OUTLIER = 0, 1, 2


class VoError(RuntimeError):
    pass


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    """pointer of a numpy array / raw int device pointer / None"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return _vp(int(a))
    return a.ctypes.data_as(_vp)


def version():
    return _L.vo_version()


def device_count():
    n = C.c_int(0)
    _L.vo_device_count(C.byref(n))
    return n.value


def pose_inverse(T):
    T = _f32(T).reshape(12)
    out = np.zeros(12, np.float32)
    _L.vo_pose_inverse(_p(T), _p(out))
    return out.reshape(3, 4)


def pose_mul(A, B):
    A = _f32(A).reshape(12)
    B = _f32(B).reshape(12)
    out = np.zeros(12, np.float32)
    _L.vo_pose_mul(_p(A), _p(B), _p(out))
    return out.reshape(3, 4)


def comm_unique_id():
    buf = np.zeros(128, np.uint8)
    st = _L.vo_comm_unique_id(_p(buf))
    if st:
        raise VoError("vo_comm_unique_id: " + _L.vo_status_str(st).decode())
    return buf


class Context:
    """vo_ctx: one GPU + one stream. `stream` may be a raw cudaStream_t (e.g. torch's)."""

    def __init__(self, device=0, stream=None):
        h = _vp()
        st = _L.vo_ctx_create(int(device), _vp(stream) if stream else None, C.byref(h))
        if st:
            raise VoError(f"vo_ctx_create(device={device}): {_L.vo_status_str(st).decode()} "
                          "(no usable CUDA device; there is no CPU fallback)")
        self._h = h

    def close(self):
        if self._h:
            _L.vo_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, what):
        if st:
            raise VoError(f"{what}: {_L.vo_status_str(st).decode()}: {_L.vo_last_error(self._h).decode()}")

    def sync(self):
        self._check(_L.vo_ctx_sync(self._h), "vo_ctx_sync")

    @property
    def kernel_launches(self):
        return _L.vo_ctx_kernel_launches(self._h)

    @property
    def stream(self):
        return _L.vo_ctx_stream(self._h)

    def comm_init(self, n_ranks, rank, unique_id):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        self._check(_L.vo_ctx_comm_init(self._h, n_ranks, rank, _p(uid)), "vo_ctx_comm_init")

    def comm_destroy(self):
        _L.vo_ctx_comm_destroy(self._h)

    # fused exchange over NVLink peer memory: export my mailbox handle, attach everybody's
    def peer_export(self):
        buf = np.zeros(64, np.uint8)
        self._check(_L.vo_ctx_peer_export(self._h, _p(buf)), "vo_ctx_peer_export")
        return buf

    def peer_attach(self, n_ranks, rank, handles):
        h = np.ascontiguousarray(handles, np.uint8).reshape(n_ranks * 64)
        self._check(_L.vo_ctx_peer_attach(self._h, n_ranks, rank, _p(h)), "vo_ctx_peer_attach")

    def peer_detach(self):
        _L.vo_ctx_peer_detach(self._h)

    @property
    def peer_active(self):
        return bool(_L.vo_ctx_peer_active(self._h))

    def picp(self):
        return Picp(self)

    def selftest_reciprocal(self):
        """all 2^32 inputs of the reciprocal shortcut vs __frcp_rn: (inputs in the gate, packed mismatches,
        scalar mismatches, first bad bit pattern + 1)"""
        out = np.zeros(4, np.uint64)
        self._check(_L.vo_selftest_reciprocal(self._h, _p(out)), "vo_selftest_reciprocal")
        return tuple(int(x) for x in out)

    # ---- pr::Camera
    def project_points(self, K, rows, cols, pose, world, keep_indices=False):
        world = _f32(world).reshape(-1, 3)
        out = np.empty((max(len(world), 1), 2), np.float32)
        n_out, n_in = _i64(0), _i64(0)
        self._check(_L.vo_project_points(self._h, _p(_f32(K).reshape(9)), rows, cols, _p(_f32(pose).reshape(12)),
                                         _p(world), len(world), int(keep_indices), _p(out), C.byref(n_out),
                                         C.byref(n_in)), "vo_project_points")
        return out[: n_out.value].copy(), n_in.value

    # ---- match_points
    MATCH_AUTO, MATCH_BRUTE, MATCH_ORDERED, MATCH_INDEXED_EXACT, MATCH_INDEXED_FILTERED = 0, 1, 2, 3, 4

    def match_set_path(self, path):
        self._check(_L.vo_match_set_path(self._h, int(path)), "vo_match_set_path")

    def match(self, descA, descB, dist_thr=0.2, ratio_thr=0.8, idA=None, idB=None, row_begin=0, row_end=None):
        descA = _f32(descA)
        descB = _f32(descB)
        n1, dim = descA.shape
        n2 = descB.shape[0]
        if row_end is None:
            row_end = n1
        cap = max(row_end - row_begin, 1)
        pairs = np.zeros((cap, 2), np.int32)
        n_out = _i64(0)
        stats = np.zeros(2, np.int64)
        ia = _i32(idA) if idA is not None else None
        ib = _i32(idB) if idB is not None else None
        self._check(_L.vo_match(self._h, _p(descA), n1, _p(descB), n2, dim, dist_thr, ratio_thr, _p(ia), _p(ib),
                                row_begin, row_end, _p(pairs), cap, C.byref(n_out), _p(stats)), "vo_match")
        return pairs[: n_out.value].copy(), (int(stats[0]), int(stats[1]))

    def match_dev(self, d_descA, n1, d_descB, n2, dim, d_pairs_out, capacity, dist_thr=0.2, ratio_thr=0.8,
                  d_idA=None, d_idB=None, row_begin=0, row_end=None, d_best=None, d_second=None, d_idx=None):
        """all d_* are raw device pointers (ints). Returns (n_matches, (possible, correct))."""
        if row_end is None:
            row_end = n1
        n_out = _i64(0)
        stats = np.zeros(2, np.int64)
        self._check(_L.vo_match_dev(self._h, _p(d_descA), n1, _p(d_descB), n2, dim, dist_thr, ratio_thr,
                                    _p(d_idA), _p(d_idB), row_begin, row_end, _p(d_pairs_out), capacity,
                                    C.byref(n_out), _p(stats), _p(d_best), _p(d_second), _p(d_idx)),
                    "vo_match_dev")
        return n_out.value, (int(stats[0]), int(stats[1]))

    def match_sharded_dev(self, d_descA, n1, d_descB, n2, dim, shard, n_shards, d_match_idx, d_pairs_out=None, capacity=0,
                          dist_thr=0.2, ratio_thr=0.8):
        """this shard's share of all n1 rows (Morton-order segments on the indexed path), the MAX all-reduce of the
        per-row results when a communicator is attached, and the compacted pairs; returns the pair count"""
        n_out = _i64(0)
        self._check(_L.vo_match_sharded_dev(self._h, _p(d_descA), n1, _p(d_descB), n2, dim, dist_thr, ratio_thr, shard,
                                            n_shards, _p(d_match_idx), _p(d_pairs_out), capacity, C.byref(n_out)),
                    "vo_match_sharded_dev")
        return n_out.value

    def match_compact_dev(self, d_match_idx, n1, d_pairs_out, capacity):
        n_out = _i64(0)
        self._check(_L.vo_match_compact_dev(self._h, _p(d_match_idx), n1, _p(d_pairs_out), capacity, C.byref(n_out)),
                    "vo_match_compact_dev")
        return n_out.value

    # ---- Cam
    def triangulate(self, K, T1, T2, x1, x2):
        x1 = _f32(x1).reshape(-1, 2)
        x2 = _f32(x2).reshape(-1, 2)
        out = np.zeros((len(x1), 3), np.float32)
        if len(x1):
            self._check(_L.vo_triangulate(self._h, _p(_f32(K).reshape(9)), _p(_f32(T1).reshape(12)),
                                          _p(_f32(T2).reshape(12)), _p(x1), _p(x2), len(x1), _p(out)),
                        "vo_triangulate")
        return out

    def triangulate_dev(self, K, T1, T2, d_x1, d_x2, n, d_out):
        self._check(_L.vo_triangulate_dev(self._h, _p(_f32(K).reshape(9)), _p(_f32(T1).reshape(12)),
                                          _p(_f32(T2).reshape(12)), _p(d_x1), _p(d_x2), n, _p(d_out)),
                    "vo_triangulate_dev")

    def essential_recover(self, K, x1, x2, method="ransac", prob=0.999, threshold=1.0, max_iters=1000, full=False):
        """src/cam.cpp:37-91. method "ransac": cv::findEssentialMat(RANSAC) restated (the reference's call);
        "8pt": normalised linear estimator on all matches. full=True also returns (ransac_mask, inliers, iterations)."""
        x1 = _f32(x1).reshape(-1, 2)
        x2 = _f32(x2).reshape(-1, 2)
        E = np.zeros(9)
        R = np.zeros(9)
        t = np.zeros(3)
        mask = np.zeros(max(len(x1), 1), np.uint8)
        rmask = np.zeros(max(len(x1), 1), np.uint8)
        good, rin, rit = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(_L.vo_essential_recover_ex(self._h, _p(_f32(K).reshape(9)), _p(x1), _p(x2), len(x1),
                                               0 if method == "ransac" else 1, prob, threshold, max_iters, _p(E), _p(R),
                                               _p(t), _p(mask), C.byref(good), _p(rmask), C.byref(rin), C.byref(rit)),
                    "vo_essential_recover")
        out = (E.reshape(3, 3), R.reshape(3, 3), t, mask[: len(x1)], good.value)
        return out + (rmask[: len(x1)], rin.value, rit.value) if full else out

    # ---- batched independent sequences (BASELINE config 5)
    def seq_batch_run(self, params, cnt, uv, desc, id_real, world_cap=1024):
        """host arrays cnt[S,F], uv[S,F,P,2], desc[S,F,P,10], id_real[S,F,P] -> dict of host results"""
        cnt = _i32(cnt)
        S, F = cnt.shape
        P = uv.shape[2]
        uv, desc, id_real = _f32(uv), _f32(desc), _i32(id_real)
        out = dict(poses=np.zeros((S, F, 3, 4), np.float32), world_xyz=np.zeros((S, world_cap, 3), np.float32),
                   world_id=np.zeros((S, world_cap), np.int32), world_cnt=np.zeros(S, np.int32),
                   rounds=np.zeros((S, F), np.int32), inliers=np.zeros((S, F, 2), np.int32), status=np.zeros(S, np.int32))
        self._check(_L.vo_seq_batch_run(self._h, C.byref(params), S, F, P, world_cap, _p(cnt), _p(uv), _p(desc),
                                        _p(id_real), _p(out["poses"]), _p(out["world_xyz"]), _p(out["world_id"]),
                                        _p(out["world_cnt"]), _p(out["rounds"]), _p(out["inliers"]), _p(out["status"])),
                    "vo_seq_batch_run")
        return out

    def seq_batch_run_dev(self, params, S, F, P, world_cap, d_cnt, d_uv, d_desc, d_id, d_poses, d_wxyz, d_wid, d_wcnt,
                          d_rounds=None, d_inliers=None, d_status=None):
        """raw device pointers; only enqueues on the context's stream"""
        self._check(_L.vo_seq_batch_run_dev(self._h, C.byref(params), S, F, P, world_cap, _p(d_cnt), _p(d_uv), _p(d_desc),
                                            _p(d_id), _p(d_poses), _p(d_wxyz), _p(d_wid), _p(d_wcnt), _p(d_rounds),
                                            _p(d_inliers), _p(d_status)), "vo_seq_batch_run_dev")

    def anti_join(self, matched_id, cand_id):
        m = _i32(matched_id).ravel()
        c = _i32(cand_id).ravel()
        keep = np.zeros(max(len(c), 1), np.uint8)
        n_keep = _i64(0)
        self._check(_L.vo_anti_join(self._h, _p(m) if len(m) else None, len(m), _p(c) if len(c) else None, len(c),
                                    _p(keep), C.byref(n_keep)), "vo_anti_join")
        return keep[: len(c)].astype(bool)


class Picp:
    """vo_picp: device-resident pr::PICPSolver."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = _vp()
        ctx._check(_L.vo_picp_create(ctx._h, C.byref(h)), "vo_picp_create")
        self._h = h
        self._keep = []

    def close(self):
        if self._h:
            _L.vo_picp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_camera(self, K, rows, cols, pose):
        self.ctx._check(_L.vo_picp_set_camera(self._h, _p(_f32(K).reshape(9)), rows, cols,
                                              _p(_f32(pose).reshape(12))), "vo_picp_set_camera")

    def set_pose(self, pose):
        self.ctx._check(_L.vo_picp_set_pose(self._h, _p(_f32(pose).reshape(12))), "vo_picp_set_pose")

    def get_pose(self):
        out = np.zeros(12, np.float32)
        self.ctx._check(_L.vo_picp_get_pose(self._h, _p(out)), "vo_picp_get_pose")
        return out.reshape(3, 4)

    def set_points(self, world, image):
        world = _f32(world).reshape(-1, 3)
        image = _f32(image).reshape(-1, 2)
        self.ctx._check(_L.vo_picp_set_points(self._h, _p(world), len(world), _p(image), len(image)),
                        "vo_picp_set_points")

    def set_points_ptr(self, world_ptr, n_world, image_ptr, n_image):
        """HOST pointers (e.g. pinned torch tensors): copies to the device."""
        self.ctx._check(_L.vo_picp_set_points(self._h, _vp(world_ptr), n_world, _vp(image_ptr), n_image),
                        "vo_picp_set_points")

    def set_points_dev(self, d_world, n_world, d_image, n_image):
        self.ctx._check(_L.vo_picp_set_points_dev(self._h, _vp(d_world), n_world, _vp(d_image), n_image),
                        "vo_picp_set_points_dev")

    def set_correspondences(self, pairs):
        pairs = _i32(pairs).reshape(-1, 2)
        self.ctx._check(_L.vo_picp_set_correspondences(self._h, _p(pairs) if len(pairs) else None, len(pairs)),
                        "vo_picp_set_correspondences")

    def set_correspondences_ptr(self, pairs_ptr, n):
        self.ctx._check(_L.vo_picp_set_correspondences(self._h, _vp(pairs_ptr), n), "vo_picp_set_correspondences")

    def set_correspondences_dev(self, d_pairs, n):
        self.ctx._check(_L.vo_picp_set_correspondences_dev(self._h, _vp(d_pairs), n),
                        "vo_picp_set_correspondences_dev")

    def pack(self):
        self.ctx._check(_L.vo_picp_pack(self._h), "vo_picp_pack")

    def set_mode(self, mode):
        """MODE_AUTO / MODE_STREAM (one launch per round over the packed planes) / MODE_RESIDENT (one persistent
        launch per solve, correspondences resident in shared memory)"""
        self.ctx._check(_L.vo_picp_set_mode(self._h, int(mode)), "vo_picp_set_mode")

    @property
    def resident_capacity(self):
        n = _i64(0)
        self.ctx._check(_L.vo_picp_resident_capacity(self._h, C.byref(n)), "vo_picp_resident_capacity")
        return n.value

    def linearize(self, thr, keep_outliers=False, want_status=False, n_pairs=None):
        H = np.zeros(36, np.float32)
        b = np.zeros(6, np.float32)
        st = Stats()
        status = np.zeros(max(n_pairs or 0, 1), np.uint8) if want_status else None
        self.ctx._check(_L.vo_picp_linearize(self._h, thr, int(keep_outliers), _p(H), _p(b), C.byref(st),
                                             _p(status)), "vo_picp_linearize")
        return dict(H=H.reshape(6, 6), b=b, chi_in=st.chi_inliers, chi_out=st.chi_outliers,
                    n_inliers=st.num_inliers, n_outliers=st.num_outliers,
                    status=status[: n_pairs] if want_status else None)

    def one_round(self, thr, damping=1.0, keep_outliers=False):
        st = Stats()
        self.ctx._check(_L.vo_picp_one_round(self._h, thr, damping, int(keep_outliers), C.byref(st)),
                        "vo_picp_one_round")
        return st

    def enqueue_rounds(self, thr, damping, keep_outliers, n_rounds):
        self.ctx._check(_L.vo_picp_enqueue_rounds(self._h, thr, damping, int(keep_outliers), n_rounds),
                        "vo_picp_enqueue_rounds")

    def fetch_stats(self, n_rounds):
        arr = (Stats * max(n_rounds, 1))()
        self.ctx._check(_L.vo_picp_fetch_stats(self._h, C.cast(arr, _vp), n_rounds), "vo_picp_fetch_stats")
        return list(arr)[:n_rounds]

    def solve(self, thr, damping=1.0, keep_outliers=False, max_rounds=50, rel_tol=1e-5):
        st = Stats()
        done = C.c_int(0)
        self.ctx._check(_L.vo_picp_solve(self._h, thr, damping, int(keep_outliers), max_rounds, rel_tol,
                                         C.byref(done), C.byref(st)), "vo_picp_solve")
        return done.value, st
