"""TEST INFRASTRUCTURE ONLY -- a CPU module with the name-level contract of the reference's pybind11 module
`point_cloud_lib_ops` (custom_ops/ops_list.cpp:19-26), so that the reference's own Python (custom_ops/*.py, pc/*.py,
layers/*.py, models/*.py) runs on CPU tensors in the build container to produce golden fixtures
(tests/golden/gen_fpn_golden.py).  Integer ops go through the C oracle (oracle/se3_oracle.c, pinned to the outputs of
the unmodified reference CUDA ops, tests/test_oracle_ref_ops.py); the aggregation is its scatter formulation
(feat_basis_proj.cu:55-118 / feat_basis_proj_grads.cu:100-145).  Never imported by se3conv3d_b200/."""
import numpy as np
import torch

from oracle import int_oracle as io


def _np(t, dt):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(dt))


def compute_keys(pts, batch_ids, aabb_min, grid_size, cell_size):
    return torch.from_numpy(io.compute_keys(_np(pts, np.float32), _np(batch_ids, np.int32), _np(aabb_min, np.float32),
                                            _np(grid_size, np.int32), _np(cell_size, np.float32)))


def ball_query(src, dst, batch_src, batch_dst, min_pt, num_cells, radius, max_neighbors):
    assert int(max_neighbors) == 0, "every call site of the path passes max_neighbors = 0"
    nb, ends = io.ball_query(_np(src, np.float32), _np(dst, np.float32), _np(batch_src, np.int32), _np(batch_dst, np.int32),
                             _np(min_pt, np.float32), _np(num_cells, np.int32), _np(radius, np.float32))
    return [torch.from_numpy(nb), torch.from_numpy(ends)]


def knn_query(pts, batch_ids, k):
    idx, _ = io.knn_query(_np(pts, np.float32), _np(batch_ids, np.int32), int(k))
    return torch.from_numpy(idx)


def feat_basis_proj(basis, feats, neighbors, ends):
    t = torch.zeros((ends.shape[0], feats.shape[1], basis.shape[1]), dtype=feats.dtype)
    return t.index_add(0, neighbors[:, 0].long(), feats[neighbors[:, 1].long()][:, :, None] * basis.to(feats.dtype)[:, None, :])


def feat_basis_proj_grad(basis, feats, neighbors, ends, grads):
    r, s = neighbors[:, 0].long(), neighbors[:, 1].long()
    g = grads[r]                                                    # [E', C, K]
    feat_grads = torch.zeros_like(feats).index_add(0, s, torch.einsum("eck,ek->ec", g, basis.to(grads.dtype)))
    basis_grads = torch.einsum("eck,ec->ek", g, feats)
    return [feat_grads, basis_grads]
