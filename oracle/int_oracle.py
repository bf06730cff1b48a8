"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/se3_oracle.c (numpy in / numpy out)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libse3_oracle.so")
_lib = None


def build():
    src = os.path.join(HERE, "se3_oracle.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "_build/libse3_oracle.so"], stdout=subprocess.DEVNULL)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_ball_query.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def grid_setup_ball_query(src, batch_src, radius):
    """min_pt / num_cells exactly as the reference wrapper computes them
    (custom_ops/BallQuery.py:36-41), in float32 numpy."""
    b = int(batch_src.max()) + 1
    mn = np.stack([src[batch_src == i].min(0) for i in range(b)]).astype(np.float32) - np.float32(1e-6)
    mx = np.stack([src[batch_src == i].max(0) for i in range(b)]).astype(np.float32) - np.float32(1e-6)
    nc = ((mx - mn) / np.float32(radius)).astype(np.int32) + 1
    return mn.astype(np.float32), nc.max(0).astype(np.int32)


def compute_keys(pts, batch, aabb_min, num_cells, cell_size):
    pts = np.ascontiguousarray(pts, np.float32)
    batch = np.ascontiguousarray(batch, np.int32)
    out = np.empty(pts.shape[0], np.int64)
    lib().oracle_compute_keys(_p(pts), _p(batch), C.c_int64(pts.shape[0]), _p(np.ascontiguousarray(aabb_min, np.float32)),
                              _p(np.ascontiguousarray(num_cells, np.int32)),
                              _p(np.ascontiguousarray(cell_size, np.float32)), _p(out))
    return out


def ball_query(src, dst, bsrc, bdst, min_pt, num_cells, radius3):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.ascontiguousarray(dst, np.float32)
    bsrc = np.ascontiguousarray(bsrc, np.int32)
    bdst = np.ascontiguousarray(bdst, np.int32)
    mn = np.ascontiguousarray(min_pt, np.float32)
    nc = np.ascontiguousarray(num_cells, np.int32)
    r = np.ascontiguousarray(radius3, np.float32)
    ends = np.empty(dst.shape[0], np.int32)
    args = (_p(src), _p(dst), _p(bsrc), _p(bdst), C.c_int64(src.shape[0]), C.c_int64(dst.shape[0]), _p(mn), _p(nc), _p(r))
    e = lib().oracle_ball_query(*args, None, _p(ends))
    nb = np.empty((e, 2), np.int64)
    lib().oracle_ball_query(*args, _p(nb), _p(ends))
    return nb, ends


def knn_query(pts, batch, k):
    pts = np.ascontiguousarray(pts, np.float32)
    batch = np.ascontiguousarray(batch, np.int32)
    out = np.empty((pts.shape[0], k), np.int32)
    dist = np.empty((pts.shape[0], k), np.float32)
    lib().oracle_knn_query(_p(pts), _p(batch), C.c_int64(pts.shape[0]), C.c_int32(k), _p(out), _p(dist))
    return out, dist


def canonical_rows(neighbors, ends):
    """Sort every CSR row by source index (the reference's in-row order is an atomic race)."""
    nb = np.asarray(neighbors).copy()
    order = np.lexsort((nb[:, 1], nb[:, 0]))
    return nb[order]
