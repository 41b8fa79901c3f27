"""CPU-only: the C-ABI library loads and exports every symbol include/vo_b200.h declares.
No compute calls (no GPU here); creating a context without a device must fail loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "02-visualodometry_b200", "libvo_b200.so")
HDR = os.path.join(ROOT, "include", "vo_b200.h")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vo_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("vo_picp_one_round", "vo_picp_linearize", "vo_match", "vo_triangulate", "vo_essential_recover",
                 "vo_project_points", "vo_anti_join", "vo_ctx_comm_init"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_status_strings():
    lib = ctypes.CDLL(LIB)
    lib.vo_status_str.restype = ctypes.c_char_p
    assert lib.vo_version() == 100
    assert lib.vo_status_str(0) == b"ok"
    assert b"CUDA" in lib.vo_status_str(2)


def test_pose_helpers_match_oracle(oracle):
    """vo_pose_inverse / vo_pose_mul are host arithmetic (Eigen Isometry3f semantics): bit-exact vs oracle."""
    import numpy as np
    from backends import product
    vo = product()
    rng = np.random.default_rng(0)
    for _ in range(50):
        A = rng.normal(size=(3, 4)).astype(np.float32)
        B = rng.normal(size=(3, 4)).astype(np.float32)
        assert np.array_equal(vo.pose_inverse(A), oracle.pose_inverse(A))
        assert np.array_equal(vo.pose_mul(A, B), oracle.pose_mul(A, B))


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of falling back."""
    from backends import product
    vo = product()
    if vo.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(vo.VoError):
        vo.Context(0)
