"""ctypes binding of the CPU ORACLE (oracle/libvo_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvo_oracle.so")

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    src = [os.path.join(_HERE, n) for n in ("vo_oracle.cpp", "five_point.cpp", "vo_oracle.h")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-s", "-C", _HERE, "libvo_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    L.vo_ref_project_point.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, f32p]
    L.vo_ref_project_point.restype = C.c_int
    L.vo_ref_project_points.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, C.c_int, C.c_int, f32p,
                                        C.POINTER(C.c_int)]
    L.vo_ref_project_points.restype = C.c_int
    L.vo_ref_error_jacobian.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p]
    L.vo_ref_error_jacobian.restype = C.c_int
    L.vo_ref_linearize.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, f32p, i32p, C.c_int64, C.c_float,
                                   C.c_int, C.c_int, f64p, f64p, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p]
    L.vo_ref_linearize.restype = None
    L.vo_ref_ldlt_solve6.argtypes = [f32p, f32p, f32p]
    L.vo_ref_ldlt_solve6.restype = None
    L.vo_ref_pose_update.argtypes = [f32p, f32p]
    L.vo_ref_pose_update.restype = None
    L.vo_ref_one_round.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, f32p, i32p, C.c_int64, C.c_float,
                                   C.c_float, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   C.POINTER(C.c_int)]
    L.vo_ref_one_round.restype = None
    L.vo_ref_one_round_mt.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, f32p, i32p, C.c_int64,
                                      C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.vo_ref_one_round_mt.restype = None
    L.vo_ref_match.argtypes = [f32p, C.c_int64, f32p, C.c_int64, C.c_int, C.c_float, C.c_float,
                               C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                               C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p]
    L.vo_ref_match.restype = C.c_int64
    L.vo_ref_triangulate.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int64, f32p]
    L.vo_ref_triangulate.restype = None
    L.vo_ref_essential_recover.argtypes = [f32p, f32p, f32p, C.c_int64, f64p, f64p, f64p, u8p]
    L.vo_ref_essential_recover.restype = C.c_int
    L.vo_ref_find_essential_ransac.argtypes = [f32p, f32p, f32p, C.c_int64, C.c_double, C.c_double, C.c_int, f64p, C.c_void_p,
                                               C.POINTER(C.c_int)]
    L.vo_ref_find_essential_ransac.restype = C.c_int
    L.vo_ref_essential_recover_ransac.argtypes = [f32p, f32p, f32p, C.c_int64, f64p, f64p, f64p, u8p]
    L.vo_ref_essential_recover_ransac.restype = C.c_int
    L.vo_ref_five_point.argtypes = [f64p, f64p, f64p]
    L.vo_ref_five_point.restype = C.c_int
    L.vo_ref_ransac_subsets.argtypes = [C.c_int, C.c_int, i32p]
    L.vo_ref_ransac_subsets.restype = None
    L.vo_ref_recover_pose.argtypes = [f64p, f32p, f32p, f32p, C.c_int64, f64p, f64p, u8p]
    L.vo_ref_recover_pose.restype = C.c_int
    L.vo_ref_anti_join.argtypes = [i32p, C.c_int64, i32p, C.c_int64, u8p]
    L.vo_ref_anti_join.restype = C.c_int64
    L.vo_ref_pose_inverse.argtypes = [f32p, f32p]
    L.vo_ref_pose_mul.argtypes = [f32p, f32p, f32p]
    L.vo_ref_num_threads.restype = C.c_int
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().vo_ref_num_threads()


def project_points(K, rows, cols, pose, world, keep_indices=False):
    world = _f32(world).reshape(-1, 3)
    out = np.empty((len(world), 2), np.float32)
    n_out = C.c_int(0)
    inside = lib().vo_ref_project_points(_f32(K).ravel(), rows, cols, _f32(pose).ravel(), world, len(world),
                                         int(keep_indices), out, C.byref(n_out))
    return out[: n_out.value].copy(), inside


def error_jacobian(K, rows, cols, pose, p, z):
    e = np.zeros(2, np.float32)
    J = np.zeros(12, np.float32)
    ok = lib().vo_ref_error_jacobian(_f32(K).ravel(), rows, cols, _f32(pose).ravel(), _f32(p), _f32(z), e, J)
    return bool(ok), e, J.reshape(2, 6)


def linearize(K, rows, cols, pose, world, image, pairs, thr, keep_outliers, accum="f32", want_status=True):
    """Returns dict(H[6,6] f64, b[6] f64, chi_in, chi_out, n_inliers, status[u8])."""
    world = _f32(world).reshape(-1, 3)
    image = _f32(image).reshape(-1, 2)
    pairs = _i32(pairs).reshape(-1, 2)
    H = np.zeros(36, np.float64)
    b = np.zeros(6, np.float64)
    ci, co, ni = C.c_double(0), C.c_double(0), C.c_int64(0)
    status = np.zeros(len(pairs), np.uint8) if want_status else None
    lib().vo_ref_linearize(_f32(K).ravel(), rows, cols, _f32(pose).ravel(), world, image, pairs, len(pairs),
                           float(thr), int(keep_outliers), 0 if accum == "f32" else 1, H, b, C.byref(ci),
                           C.byref(co), C.byref(ni), _ptr(status))
    return dict(H=H.reshape(6, 6), b=b, chi_in=ci.value, chi_out=co.value, n_inliers=ni.value, status=status)


def ldlt_solve6(A, rhs):
    x = np.zeros(6, np.float32)
    lib().vo_ref_ldlt_solve6(_f32(A).ravel(), _f32(rhs), x)
    return x


def pose_update(dx, pose):
    p = _f32(pose).ravel().copy()
    lib().vo_ref_pose_update(_f32(dx), p)
    return p.reshape(3, 4)


def one_round(K, rows, cols, pose, world, image, pairs, thr, damping=1.0, keep_outliers=False, n_threads=1):
    """One Gauss-Newton round; returns (new_pose[3,4], chi_in, chi_out, n_inliers)."""
    world = _f32(world).reshape(-1, 3)
    image = _f32(image).reshape(-1, 2)
    pairs = _i32(pairs).reshape(-1, 2)
    p = _f32(pose).ravel().copy()
    ci, co, ni = C.c_float(0), C.c_float(0), C.c_int(0)
    if n_threads <= 1:
        lib().vo_ref_one_round(_f32(K).ravel(), rows, cols, p, world, image, pairs, len(pairs), float(thr),
                               float(damping), int(keep_outliers), C.byref(ci), C.byref(co), C.byref(ni))
    else:
        lib().vo_ref_one_round_mt(_f32(K).ravel(), rows, cols, p, world, image, pairs, len(pairs), float(thr),
                                  float(damping), int(keep_outliers), int(n_threads), C.byref(ci),
                                  C.byref(co), C.byref(ni))
    return p.reshape(3, 4), ci.value, co.value, ni.value


def match(descA, descB, dist_thr=0.2, ratio_thr=0.8, idA=None, idB=None, row_begin=0, row_end=None,
          order="eigen", n_threads=1, want_rows=False):
    """Brute-force matcher. Returns (pairs[int32 n,2], stats(possible, correct)[, best, second, idx])."""
    descA = _f32(descA)
    descB = _f32(descB)
    n1, dim = descA.shape if descA.ndim == 2 else (0, descB.shape[1] if descB.ndim == 2 else 0)
    n2 = descB.shape[0] if descB.ndim == 2 else 0
    if row_end is None:
        row_end = n1
    rows = row_end - row_begin
    pairs = np.zeros((max(rows, 1), 2), np.int32)
    stats = (C.c_int64 * 2)(0, 0)
    ia = _i32(idA) if idA is not None else None
    ib = _i32(idB) if idB is not None else None
    best = np.zeros(max(rows, 1), np.float32) if want_rows else None
    second = np.zeros(max(rows, 1), np.float32) if want_rows else None
    idx = np.zeros(max(rows, 1), np.int32) if want_rows else None
    n = lib().vo_ref_match(descA.reshape(-1) if descA.size else np.zeros(1, np.float32), n1,
                           descB.reshape(-1) if descB.size else np.zeros(1, np.float32), n2, dim,
                           float(dist_thr), float(ratio_thr), _ptr(ia), _ptr(ib), row_begin, row_end,
                           0 if order == "eigen" else 1, int(n_threads), _ptr(pairs), stats, _ptr(best),
                           _ptr(second), _ptr(idx))
    out = (pairs[:n].copy(), (stats[0], stats[1]))
    if want_rows:
        out = out + (best[:rows], second[:rows], idx[:rows])
    return out


def triangulate(K, T1, T2, x1, x2):
    x1 = _f32(x1).reshape(-1, 2)
    x2 = _f32(x2).reshape(-1, 2)
    out = np.zeros((len(x1), 3), np.float32)
    if len(x1):
        lib().vo_ref_triangulate(_f32(K).ravel(), _f32(T1).ravel(), _f32(T2).ravel(), x1, x2, len(x1), out)
    return out


def essential_recover(K, x1, x2, method="ransac"):
    """src/cam.cpp:37-91. method "ransac": cv::findEssentialMat(RANSAC) restated (what the reference runs);
    "8pt": the normalised linear estimator on all matches (the batched-sequence option)."""
    x1 = _f32(x1).reshape(-1, 2)
    x2 = _f32(x2).reshape(-1, 2)
    E = np.zeros(9)
    R = np.zeros(9)
    t = np.zeros(3)
    mask = np.zeros(max(len(x1), 1), np.uint8)
    fn = lib().vo_ref_essential_recover_ransac if method == "ransac" else lib().vo_ref_essential_recover
    good = fn(_f32(K).ravel(), x1, x2, len(x1), E, R, t, mask)
    return E.reshape(3, 3), R.reshape(3, 3), t, mask[: len(x1)], good


def find_essential_ransac(K, x1, x2, prob=0.999, threshold=1.0, max_iters=1000):
    """cv::findEssentialMat(x1, x2, K, RANSAC, prob, threshold, maxIters) restated -> (E, mask, good, iterations)"""
    x1 = _f32(x1).reshape(-1, 2)
    x2 = _f32(x2).reshape(-1, 2)
    E = np.zeros(9)
    mask = np.zeros(max(len(x1), 1), np.uint8)
    iters = C.c_int(0)
    good = lib().vo_ref_find_essential_ransac(_f32(K).ravel(), x1, x2, len(x1), prob, threshold, max_iters, E,
                                              mask.ctypes.data_as(C.c_void_p), C.byref(iters))
    return E.reshape(3, 3), mask[: len(x1)], good, iters.value


def five_point(q1, q2):
    q1 = np.ascontiguousarray(q1, np.float64).reshape(10)
    q2 = np.ascontiguousarray(q2, np.float64).reshape(10)
    out = np.zeros(90)
    n = lib().vo_ref_five_point(q1, q2, out)
    return out[: 9 * n].reshape(n, 3, 3)


def ransac_subsets(n, n_subsets):
    out = np.zeros((n_subsets, 5), np.int32)
    lib().vo_ref_ransac_subsets(int(n), int(n_subsets), out)
    return out


def recover_pose(E, K, x1, x2):
    x1 = _f32(x1).reshape(-1, 2)
    x2 = _f32(x2).reshape(-1, 2)
    R = np.zeros(9)
    t = np.zeros(3)
    mask = np.zeros(max(len(x1), 1), np.uint8)
    good = lib().vo_ref_recover_pose(np.ascontiguousarray(E, np.float64).ravel(), _f32(K).ravel(), x1, x2,
                                     len(x1), R, t, mask)
    return R.reshape(3, 3), t, mask[: len(x1)], good


def anti_join(matched_id_meas, cand_second_id_meas):
    m = _i32(matched_id_meas).ravel()
    c = _i32(cand_second_id_meas).ravel()
    keep = np.zeros(max(len(c), 1), np.uint8)
    lib().vo_ref_anti_join(m if len(m) else np.zeros(1, np.int32), len(m),
                           c if len(c) else np.zeros(1, np.int32), len(c), keep)
    return keep[: len(c)].astype(bool)


def pose_inverse(T):
    out = np.zeros(12, np.float32)
    lib().vo_ref_pose_inverse(_f32(T).ravel(), out)
    return out.reshape(3, 4)


def pose_mul(A, B):
    out = np.zeros(12, np.float32)
    lib().vo_ref_pose_mul(_f32(A).ravel(), _f32(B).ravel(), out)
    return out.reshape(3, 4)
