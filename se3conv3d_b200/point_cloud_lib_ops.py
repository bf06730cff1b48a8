"""Drop-in replacement of the reference's pybind11 module `point_cloud_lib_ops`
(point_cloud_lib/custom_ops/ops_list.cpp:19-26): same five entry points, same tensor contracts,
backed by the C ABI of libse3conv3d_b200.so.

    feat_basis_proj(basis f32 [E,K], feats f32 [N,C], neighbors i32 [E,2], ends i32 [M]) -> [M,C,K]
    feat_basis_proj_grad(basis, feats, neighbors, ends, grads [M,C,K]) -> [feat_grads [N,C], basis_grads [E,K]]
    ball_query(src [N,3], dst [M,3], batch_src i32, batch_dst i32, min_pt [B,3], num_cells i32 [3],
               radius f32 [3], max_neighbors) -> [neighbors i64 [E,2], ends i32 [M]]
    knn_query(pts f32 [N,3], batch i32 [N], k) -> i32 [N,k]
    compute_keys(pts f32 [N,3], batch i32 [N], aabb_min f32 [B,3], grid_size i32 [3], cell_size f32 [3]) -> i64 [N]

Unlike the reference (which returns zeros for unsupported K / D, feat_basis_utils.cuh:35-41) every
unsupported argument raises.
"""
import torch

from ._lib import lib, check, ptr, stream, workspace, Se3Error


def _f32(t):
    return t.to(torch.float32).contiguous()


def _i32(t):
    return t.to(torch.int32).contiguous()


def compute_keys(p_pts, p_batch_ids, p_aabb_min, p_grid_size, p_cell_size):
    if p_pts.dim() != 2 or p_pts.shape[1] != 3:
        raise Se3Error("compute_keys: only 3-D point clouds are supported")
    pts, b = _f32(p_pts), _i32(p_batch_ids)
    amin, gs, cs = _f32(p_aabb_min), _i32(p_grid_size), _f32(p_cell_size)
    out = torch.empty(pts.shape[0], dtype=torch.int64, device=pts.device)
    check(lib().se3_compute_keys(ptr(pts), ptr(b), pts.shape[0], ptr(amin), ptr(gs), ptr(cs), ptr(out), stream()),
          "se3_compute_keys")
    return out


def ball_query(p_pt_src, p_pt_dest, p_batch_ids_src, p_batch_ids_dest, p_min_pt, p_num_cells, p_radius,
               p_max_neighbors=0):
    if p_max_neighbors not in (0, None):
        # dead in the reference on this path: every call site passes 0 (pc/BQNeighborhood.py:20)
        raise Se3Error("ball_query: max_neighbors > 0 (random neighbour cap) is not supported")
    if p_pt_src.shape[1] != 3:
        raise Se3Error("ball_query: only 3-D point clouds are supported")
    src, dst = _f32(p_pt_src), _f32(p_pt_dest)
    bs, bd = _i32(p_batch_ids_src), _i32(p_batch_ids_dest)
    mn, nc, rad = _f32(p_min_pt), _i32(p_num_cells), _f32(p_radius)
    n, m = src.shape[0], dst.shape[0]
    L = lib()
    ws_bytes = L.se3_ball_query_workspace_bytes(n, m)
    ws = workspace(ws_bytes, src.device, 'bq')
    ends = torch.empty(m, dtype=torch.int32, device=src.device)
    total = torch.empty(1, dtype=torch.int64, device=src.device)
    check(L.se3_ball_query_count(ptr(src), ptr(dst), ptr(bs), ptr(bd), n, m, ptr(mn), ptr(nc), ptr(rad), ptr(ws),
                                 ws.numel(), ptr(ends), ptr(total), stream()), "se3_ball_query_count")
    e = int(total.item())  # the one host sync: the caller owns the output allocation
    neighbors = torch.empty((e, 2), dtype=torch.int64, device=src.device)
    check(L.se3_ball_query_fill(ptr(dst), n, m, ptr(rad), ptr(ws), ws.numel(), ptr(ends), e, ptr(neighbors),
                                stream()), "se3_ball_query_fill")
    return [neighbors, ends]


def knn_query(p_pt_src, p_batch_ids_src, p_k):
    if p_pt_src.shape[1] != 3:
        raise Se3Error("knn_query: only 3-D point clouds are supported")
    pts, b = _f32(p_pt_src), _i32(p_batch_ids_src)
    n = pts.shape[0]
    L = lib()
    ws = workspace(L.se3_knn_workspace_bytes(n), pts.device, 'knn')
    out = torch.empty((n, int(p_k)), dtype=torch.int32, device=pts.device)
    check(L.se3_knn_query(ptr(pts), ptr(b), n, int(p_k), ptr(ws), ws.numel(), ptr(out), stream()), "se3_knn_query")
    return out


def feat_basis_proj(p_pt_basis, p_pt_features, p_neighbors, p_start_ids):
    basis, feats = _f32(p_pt_basis), _f32(p_pt_features)
    nb, ends = _i32(p_neighbors), _i32(p_start_ids)
    m, c, k = ends.shape[0], feats.shape[1], basis.shape[1]
    out = torch.empty((m, c, k), dtype=torch.float32, device=feats.device)
    check(lib().se3_feat_basis_proj(ptr(basis), ptr(feats), ptr(nb), ptr(ends), nb.shape[0], m, c, k, ptr(out),
                                    stream()), "se3_feat_basis_proj")
    return out


def feat_basis_proj_grad(p_pt_basis, p_pt_features, p_neighbors, p_start_ids, p_grads):
    basis, feats = _f32(p_pt_basis), _f32(p_pt_features)
    nb, ends, grads = _i32(p_neighbors), _i32(p_start_ids), _f32(p_grads)
    m, c, k = ends.shape[0], feats.shape[1], basis.shape[1]
    fg = torch.empty_like(feats)
    bg = torch.empty_like(basis)
    check(lib().se3_feat_basis_proj_grad(ptr(basis), ptr(feats), ptr(nb), ptr(ends), ptr(grads), nb.shape[0], m,
                                         feats.shape[0], c, k, ptr(fg), ptr(bg), stream()),
          "se3_feat_basis_proj_grad")
    return [fg, bg]
