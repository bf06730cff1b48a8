"""Rotation / reference-frame helpers with the reference's names (pc/RotationFunctions.py).
The per-point work (PCA frames, quaternion -> matrix) runs in the CUDA library; the small
closed-form helpers are plain tensor algebra kept for API compatibility and for tests."""
from itertools import product

import numpy as np
import torch

from .._lib import lib, check, ptr, stream, Se3Error


def all_index_combinations(n_A, n_B, device=None):
    """[(a, b)] for a in range(n_A), b in range(n_B) (pc/RotationFunctions.py:16-21)."""
    return torch.tensor(list(product(range(n_A), range(n_B))), device=device)


def quaternion_to_matrix(quaternions):
    """Real-part-first quaternions [...,4] -> rotation matrices [...,3,3] through se3_quat_frames
    for CUDA input (pc/RotationFunctions.py:53-82)."""
    q = quaternions.to(torch.float32).reshape(-1, 4).contiguous()
    out = torch.empty((q.shape[0], 9), dtype=torch.float32, device=q.device)
    check(lib().se3_quat_frames(ptr(q), q.shape[0], ptr(out), stream()), "se3_quat_frames")
    return out.reshape(quaternions.shape[:-1] + (3, 3))


def random_rotations(n, dtype=None, device=None):
    """n random rotation matrices from normalised Gaussian quaternions; consumes exactly one
    torch.randn((n,4)) like the reference (pc/RotationFunctions.py:176-216)."""
    o = torch.randn((n, 4), dtype=dtype, device=device)
    return quaternion_to_matrix(o)  # the kernel applies the copysign normalisation itself


def random_rotation(dtype=None, device=None):
    return random_rotations(1, dtype, device)[0]


def matrix_to_rotation_6d(matrix):
    """First two rows of a rotation matrix, flattened (pc/RotationFunctions.py:236-252)."""
    return matrix[..., :2, :].clone().reshape(matrix.shape[:-2] + (6,))


def sample_reference_frames_pca(points, p_neighborhood, axis_fixed=False, dtype=None, device=None):
    """Per-point PCA frames [N, 4 or 2, 9] from a k-NN neighbourhood (pc/RotationFunctions.py:307-406)."""
    k = p_neighborhood.k_
    n = points.shape[0]
    knn = getattr(p_neighborhood, "knn_table_", None)
    if knn is None:
        knn = p_neighborhood.neighbors_[:, 1].to(torch.int32).reshape(n, k)
    knn = knn.contiguous()
    fixed = -1 if (axis_fixed is None or axis_fixed is False or not axis_fixed) else int(axis_fixed)
    nf = 2 if fixed > 0 else 4
    pts = points.to(torch.float32).contiguous()
    out = torch.empty((n, nf, 9), dtype=torch.float32, device=pts.device)
    check(lib().se3_pca_frames(ptr(pts), ptr(knn), n, k, fixed, ptr(out), stream()), "se3_pca_frames")
    return out


def sample_global_reference_frames_pca(points, axis_fixed=False, dtype=None, device=None):
    """One PCA frame set per batch item [B,4,9] from [B,n,3] points (pc/RotationFunctions.py:265-304).
    Not used by any shipped config; runs the same kernel with every point of the item as neighbour."""
    if not (axis_fixed is None or not axis_fixed):
        raise NotImplementedError("Sampling global ref frames with fixed axes is not implemented")
    b, n, _ = points.shape
    flat = points.reshape(b * n, 3).to(torch.float32).contiguous()
    knn = (torch.arange(b, device=flat.device, dtype=torch.int32)[:, None] * n +
           torch.arange(n, device=flat.device, dtype=torch.int32)[None, :])
    # frames of the first point of every item, using all n points as its neighbourhood
    knn_full = knn.repeat_interleave(n, dim=0).contiguous()
    out = torch.empty((b * n, 4, 9), dtype=torch.float32, device=flat.device)
    check(lib().se3_pca_frames(ptr(flat), ptr(knn_full), b * n, n, -1, ptr(out), stream()), "se3_pca_frames")
    return out.reshape(b, n, 4, 9)[:, 0].contiguous()


def sample_reference_frames(n_origins, n_frames, axis_fixed=None, dtype=None, device=None):
    """Random frames [n_origins, n_frames, 9]: uniform SO(3), or a uniform angle about the fixed
    axis (pc/RotationFunctions.py:428-508)."""
    if axis_fixed is None or not axis_fixed:
        rot = random_rotations(n_origins * n_frames, dtype=dtype, device=device)
        return rot.reshape(n_origins, n_frames, 9)
    ang = torch.rand(n_origins * n_frames, device=device) * 2 * np.pi
    c, s, z, o = torch.cos(ang), torch.sin(ang), torch.zeros_like(ang), torch.ones_like(ang)
    if axis_fixed == 0:
        m = (o, z, z, z, c, -s, z, s, c)
    elif axis_fixed == 1:
        m = (c, z, s, z, o, z, -s, z, c)
    elif axis_fixed == 2:
        m = (c, -s, z, s, c, z, z, z, o)
    else:
        raise ValueError("axis_fixed must be 0, 1 or 2")
    return torch.stack(m, -1).reshape(n_origins, n_frames, 9)


def get_relative_rot(frames_A, frames_B, return_representation="matrix"):
    """R_A^T R_B for all frame pairs, pair index a*F_B + b (pc/RotationFunctions.py:549-600)."""
    if return_representation not in ("matrix", "6D", "quaternion"):
        raise ValueError("return_representation must be 'matrix', '6D' or 'quaternion'")
    n, fa = frames_A.shape[0], frames_A.shape[1]
    fb = frames_B.shape[1]
    A = frames_A.reshape(n, fa, 1, 3, 3)
    B = frames_B.reshape(n, 1, fb, 3, 3)
    rel = torch.matmul(A.transpose(-1, -2), B).reshape(n, fa * fb, 3, 3)
    if return_representation == "matrix":
        return rel.reshape(n, fa * fb, 9)
    if return_representation == "quaternion":
        return matrix_to_quaternion(rel)
    return matrix_to_rotation_6d(rel)


def matrix_to_quaternion(matrix):
    """Rotation matrices [..., 3, 3] -> quaternions [..., 4], real part first; of the four algebraically equal
    candidates the best conditioned one (largest |component|) is taken (pc/RotationFunctions.py:114-173)."""
    batch = matrix.shape[:-2]
    m = matrix.reshape(batch + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m, dim=-1)
    q_abs = torch.sqrt(torch.clamp(torch.stack((1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22,
                                                1.0 - m00 - m11 + m22), dim=-1), min=0.0))
    cand = torch.stack((torch.stack((q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01), dim=-1),
                        torch.stack((m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20), dim=-1),
                        torch.stack((m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21), dim=-1),
                        torch.stack((m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2), dim=-1)), dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp_min(0.1))
    pick = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, pick[..., None, None].expand(batch + (1, 4))).squeeze(-2)


def change_points_to_local_frame(points, origins, ref_frames):
    """R^T (p - o) for every frame of every origin (pc/RotationFunctions.py:603-634)."""
    R = ref_frames.reshape(ref_frames.shape[0], ref_frames.shape[1], 3, 3)
    d = (points - origins)[:, None, :, None]
    return torch.matmul(R.transpose(-1, -2), d).squeeze(-1)


def change_direction_to_local_frame(direction_vector, ref_frames):
    """d^T R (row vector times matrix) for every frame (pc/RotationFunctions.py:637-665)."""
    R = ref_frames.reshape(ref_frames.shape[0], ref_frames.shape[1], 3, 3)
    return torch.matmul(direction_vector[:, None, None, :], R).squeeze(2)


def random_rotate(p_hierarchy):
    """Applies one random global rotation to every level: points (row vectors) and frames
    (column axes) (pc/RotationFunctions.py:412-425)."""
    rot = random_rotation(device=p_hierarchy.pcs_[0].pts_.device)
    for pc in p_hierarchy.pcs_:
        pc.pts_ = torch.matmul(pc.pts_, rot.transpose(1, 0))
        fr = pc.local_frames_.reshape(pc.local_frames_.shape[0], pc.local_frames_.shape[1], 3, 3)
        pc.local_frames_ = torch.matmul(rot, fr).reshape(fr.shape[0], fr.shape[1], 9)
    return p_hierarchy
