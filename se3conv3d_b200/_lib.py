"""ctypes binding of libse3conv3d_b200.so (the C ABI declared in include/se3conv3d_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an exception is
raised.  Nothing in this package ever routes a product call through `oracle/` or a CPU path.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# SE3CONV3D_LIB selects an alternative build of the same library (kernel-tuning A/B runs); default = in-tree build
LIB_PATH = os.environ.get("SE3CONV3D_LIB") or os.path.join(_PKG, "lib", "libse3conv3d_b200.so")

# every symbol include/se3conv3d_b200.h declares (tests check the .so exports exactly these)
ABI_SYMBOLS = [
    "se3_abi_version", "se3_last_error", "se3_launch_count", "se3_profile_enable", "se3_profile_read",
    "se3_compute_keys", "se3_grid_setup", "se3_grid_cells_workspace_bytes", "se3_grid_cells", "se3_frames_select",
    "se3_ball_query_workspace_bytes", "se3_ball_query_count", "se3_ball_query_fill",
    "se3_csr_transpose_workspace_bytes", "se3_csr_transpose",
    "se3_knn_workspace_bytes", "se3_knn_query", "se3_knn_cross", "se3_pca_frames", "se3_quat_frames",
    "se3_segment_pool_f32",
    "se3_feat_basis_proj", "se3_feat_basis_proj_grad",
    "se3_conv_fwd_workspace_bytes", "se3_conv_bwd_workspace_bytes", "se3_conv_saved_bytes", "se3_conv_weight_cache_bytes",
    "se3_conv_fwd", "se3_conv_bwd", "se3_gemm_bf16_tn", "se3_gemm_bf16_mn", "se3_pack_records", "se3_conv_set_fused",
    "se3_gamma_skip_workspace_bytes", "se3_gamma_skip_fwd", "se3_gamma_skip_bwd", "se3_frame_pool_fwd", "se3_frame_pool_bwd",
    "se3_batch_pool_fwd", "se3_batch_pool_bwd",
    "se3_ball_query_fill_csr", "se3_csr_transpose_i32", "se3_segment_first_i32", "se3_segment_pick",
    "se3_hierarchy_build", "se3_bbox", "se3_grid_extents", "se3_ball_query_src_workspace_bytes",
    "se3_ball_query_dst_workspace_bytes", "se3_ball_query_prepare", "se3_ball_query_count_prepared",
    "se3_ball_query_fill_csr_prepared",
]


class ConvDesc(C.Structure):
    """Mirror of `struct se3_conv_desc`."""
    _fields_ = [
        ("n_in", C.c_int64), ("n_out", C.c_int64), ("n_edges", C.c_int64),
        ("f_in", C.c_int32), ("f_out", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32),
        ("k", C.c_int32), ("act", C.c_int32), ("precision", C.c_int32), ("reserved", C.c_int32),
        ("norm_neigh_dist", C.c_float), ("out_scale", C.c_float),
        ("pts_in", C.c_void_p), ("pts_out", C.c_void_p), ("frames_in", C.c_void_p), ("frames_out", C.c_void_p),
        ("row_ends", C.c_void_p), ("col_src", C.c_void_p),
        ("t_row_ends", C.c_void_p), ("t_edge", C.c_void_p), ("t_dst", C.c_void_p),
        ("rec_in", C.c_void_p), ("rec_out", C.c_void_p),
        ("proj_axes", C.c_void_p), ("proj_biases", C.c_void_p), ("conv_weights", C.c_void_p),
        ("weight_cache", C.c_void_p), ("weight_cache_state", C.c_int32), ("reserved2", C.c_int32),
    ]


HIER_MAX_CLOUDS, HIER_MAX_NEIGH = 10, 32


class HierDesc(C.Structure):
    """Mirror of `struct se3_hier_desc`."""
    _fields_ = [
        ("n", C.c_int64), ("n_batches", C.c_int32), ("n_pool", C.c_int32), ("init_cell", C.c_float),
        ("cells", C.c_float * HIER_MAX_CLOUDS), ("knn_k", C.c_int32), ("n_frames", C.c_int32),
        ("fixed_axis", C.c_int32), ("out_cloud", C.c_int32), ("n_neigh", C.c_int32),
        ("neigh_src", C.c_int32 * HIER_MAX_NEIGH), ("neigh_dst", C.c_int32 * HIER_MAX_NEIGH),
        ("neigh_radius", C.c_float * HIER_MAX_NEIGH),
    ]


class HierCloud(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n", "pts", "batch", "frames", "rec", "m", "cell_ids", "sorted_ids", "cell_ends")]


class HierNeigh(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("e", "row_ends", "col_src", "edge_dst", "t_row_ends", "t_edge", "t_dst")]


class HierResult(C.Structure):
    """Mirror of `struct se3_hier_result`."""
    _fields_ = [
        ("arena_used", C.c_int64), ("n_clouds", C.c_int32), ("reserved", C.c_int32), ("raw", HierCloud),
        ("out_picked", C.c_int64), ("clouds", HierCloud * HIER_MAX_CLOUDS), ("neigh", HierNeigh * HIER_MAX_NEIGH),
    ]


_lib = None


class LazyAttrs(object):
    """Attributes that are windows into the fused hierarchy arena are created on first access: an object built by
    pc.build_point_hierarchy carries `_lazy = {name: thunk}`; everything else behaves like a plain attribute.
    (A training step touches a handful of the ~150 tensors one hierarchy exposes; creating all of them eagerly
    cost more host time than the native build call.)"""

    def __getattr__(self, name):
        lazy = self.__dict__.get("_lazy")
        if lazy is not None and name in lazy:
            val = lazy.pop(name)()
            self.__dict__[name] = val
            return val
        raise AttributeError("%s has no attribute %r" % (type(self).__name__, name))


class Se3Error(RuntimeError):
    pass


def lib():
    """Returns the loaded library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Se3Error(
            "libse3conv3d_b200.so is missing (%s). Build it with `python -m se3conv3d_b200.build`; "
            "this package has no CPU / PyTorch fallback by design." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, sz, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t, C.c_float
    L.se3_abi_version.restype = C.c_int
    L.se3_last_error.restype = C.c_char_p
    L.se3_launch_count.restype = i64
    L.se3_profile_enable.argtypes = [i32]
    L.se3_profile_enable.restype = None
    L.se3_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.se3_compute_keys.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    L.se3_grid_setup.argtypes = [vp, vp, i64, i32, f32, f32, vp, vp, vp, vp]
    L.se3_grid_cells_workspace_bytes.argtypes = [i64]
    L.se3_grid_cells_workspace_bytes.restype = sz
    L.se3_grid_cells.argtypes = [vp, vp, i64, vp, vp, f32, vp, sz, vp, vp, vp, vp, i32, vp]
    L.se3_frames_select.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.se3_ball_query_workspace_bytes.argtypes = [i64, i64]
    L.se3_ball_query_workspace_bytes.restype = sz
    L.se3_ball_query_count.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, sz, vp, vp, vp]
    L.se3_ball_query_fill.argtypes = [vp, i64, i64, vp, vp, sz, vp, i64, vp, vp]
    L.se3_csr_transpose_workspace_bytes.argtypes = [i64, i64]
    L.se3_csr_transpose_workspace_bytes.restype = sz
    L.se3_csr_transpose.argtypes = [vp, i64, i64, i64, vp, sz, vp, vp, vp, vp, vp]
    L.se3_knn_workspace_bytes.argtypes = [i64]
    L.se3_knn_workspace_bytes.restype = sz
    L.se3_knn_query.argtypes = [vp, vp, i64, i32, vp, sz, vp, vp]
    L.se3_knn_cross.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp]
    L.se3_pca_frames.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.se3_quat_frames.argtypes = [vp, i64, vp, vp]
    L.se3_segment_pool_f32.argtypes = [vp, i64, i32, vp, vp, i64, i32, vp, vp]
    L.se3_feat_basis_proj.argtypes = [vp, vp, vp, vp, i64, i64, i32, i32, vp, vp]
    L.se3_feat_basis_proj_grad.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, vp, vp, vp]
    dp = C.POINTER(ConvDesc)
    for n in ("se3_conv_fwd_workspace_bytes", "se3_conv_bwd_workspace_bytes", "se3_conv_saved_bytes",
              "se3_conv_weight_cache_bytes"):
        getattr(L, n).argtypes = [dp]
        getattr(L, n).restype = sz
    L.se3_conv_fwd.argtypes = [dp, vp, vp, vp, vp, sz, vp]
    L.se3_conv_set_fused.argtypes = [i32]
    L.se3_gamma_skip_workspace_bytes.argtypes = [i64, i32]
    L.se3_gamma_skip_workspace_bytes.restype = sz
    L.se3_gamma_skip_fwd.argtypes = [vp, vp, vp, vp, vp, i32, i64, i32, vp, vp]
    L.se3_gamma_skip_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i64, i32, vp, vp, vp, sz, vp]
    L.se3_frame_pool_fwd.argtypes = [vp, i64, i32, i32, i32, vp, vp]
    L.se3_frame_pool_bwd.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp, vp]
    L.se3_batch_pool_fwd.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.se3_batch_pool_bwd.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
    L.se3_gemm_bf16_tn.argtypes = [vp, vp, i64, i64, i64, f32, vp, i32, i32, vp]
    L.se3_gemm_bf16_mn.argtypes = [vp, vp, i64, i64, i64, f32, vp, i32, vp]
    L.se3_pack_records.argtypes = [vp, vp, i64, i32, vp, vp]
    L.se3_ball_query_fill_csr.argtypes = [vp, i64, i64, vp, vp, sz, vp, i64, vp, vp, vp]
    L.se3_csr_transpose_i32.argtypes = [vp, vp, i64, i64, vp, sz, vp, vp, vp, vp]
    L.se3_segment_first_i32.argtypes = [vp, vp, vp, i64, vp, vp]
    L.se3_segment_pick.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
    L.se3_bbox.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    L.se3_grid_extents.argtypes = [vp, vp, i32, f32, f32, vp, vp, vp, vp]
    L.se3_ball_query_src_workspace_bytes.argtypes = [i64, i64]
    L.se3_ball_query_src_workspace_bytes.restype = sz
    L.se3_ball_query_dst_workspace_bytes.argtypes = [i64]
    L.se3_ball_query_dst_workspace_bytes.restype = sz
    L.se3_ball_query_prepare.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp, sz, i32, vp]
    L.se3_ball_query_count_prepared.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, sz, vp, sz, vp, vp, vp]
    L.se3_ball_query_fill_csr_prepared.argtypes = [vp, i64, i64, i64, vp, vp, sz, vp, sz, vp, i64, vp, vp, vp]
    L.se3_hierarchy_build.argtypes = [C.POINTER(HierDesc), vp, vp, vp, vp, vp, sz, C.POINTER(HierResult), vp]
    L.se3_conv_bwd.argtypes = [dp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    for n in ABI_SYMBOLS:
        f = getattr(L, n)
        if n not in ("se3_last_error", "se3_launch_count", "se3_profile_enable") and not n.endswith("_bytes"):
            f.restype = C.c_int
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().se3_last_error().decode("utf-8", "replace")
        raise Se3Error("%s failed (%d): %s" % (what or "se3 call", rc, msg))


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise Se3Error("se3conv3d_b200 kernels need CUDA tensors (got %s); there is no CPU fallback" % t.device)
    if not t.is_contiguous():
        raise Se3Error("tensor must be contiguous")
    return t.data_ptr()


def stream():
    """Raw handle of torch's current CUDA stream (the private fast accessor when available: the public
    torch.cuda.current_stream() builds a Stream object per call, ~10 us on the hot path)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


def grid_setup(pts, batch_ids, n_batches, cell, max_pad):
    """Per-batch (min - 1e-6, max + max_pad) and the grid extents, on the device, no host sync."""
    pts = pts.to(torch.float32).contiguous()
    b = batch_ids.to(torch.int32).contiguous()
    mn = torch.empty((n_batches, 3), dtype=torch.float32, device=pts.device)
    mx = torch.empty((n_batches, 3), dtype=torch.float32, device=pts.device)
    nc = torch.empty(3, dtype=torch.int32, device=pts.device)
    check(lib().se3_grid_setup(ptr(pts), ptr(b), pts.shape[0], int(n_batches), float(cell), float(max_pad), ptr(mn),
                               ptr(mx), ptr(nc), stream()), "se3_grid_setup")
    return mn, mx, nc


def num_batches(pc):
    """Number of batch items of a cloud as a host int, cached on the object (one sync per root cloud;
    sub-sampled clouds inherit it)."""
    nb = getattr(pc, "batch_size_host_", None)
    if nb is None:
        nb = int(pc.batch_size_)
        try:
            pc.batch_size_host_ = nb
        except AttributeError:
            pass
    return nb


_ws_pool = {}


def workspace(nbytes, device, tag=None):
    """Scratch memory for one C-ABI call.  With a `tag` the buffer is pooled per (device, stream, tag) and
    only grows: scratch of calls issued in order on one stream can be reused without a fresh allocation.
    Buffers that must outlive the call (saved-for-backward, cached CSR state) are allocated untagged."""
    nbytes = max(int(nbytes), 256)
    if torch.device(device).type != "cuda":
        raise Se3Error("se3conv3d_b200 kernels need CUDA tensors (got %s); there is no CPU fallback" % device)
    if tag is None:
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (str(device), stream(), tag)
    buf = _ws_pool.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _ws_pool[key] = buf
    return buf


PROF_KERNELS = ("k_agg_tc forward aggregation", "k_agg_tc transposed aggregation (data gradient)",
                "k_edge_tc edge gradient")


def profile_kernels(fn, reps=1):
    """Runs fn() reps times with per-kernel device timing on; returns [(name, avg ms per launch, launches)]."""
    L = lib()
    ms = (C.c_double * 3)()
    cnt = (C.c_int64 * 3)()
    L.se3_profile_read(ms, cnt)  # drain
    L.se3_profile_enable(1)
    try:
        for _ in range(reps):
            fn()
    finally:
        L.se3_profile_enable(0)
    check(L.se3_profile_read(ms, cnt), "se3_profile_read")
    return [(PROF_KERNELS[i], (ms[i] / cnt[i]) if cnt[i] else 0.0, int(cnt[i])) for i in range(3)]


def set_fused_mode(mode):
    """Selects the precision-1 kernel family (see se3_conv_set_fused in the header); returns the previous mode.
    Cached descriptor byte counts belong to the mode they were computed under: geometry objects created before
    the switch must drop their `_desc_cache`."""
    return int(lib().se3_conv_set_fused(int(mode)))


def launch_count():
    return int(lib().se3_launch_count())
