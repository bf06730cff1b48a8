"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch tensors on the host, any float dtype) of the
floating-point part of the reference hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs may import this; se3conv3d_b200/ never does.

Follows (paths relative to /root/reference/point_cloud_lib/point_cloud_lib):
  rot_tensors        layers/PNEConvLayerRotEquiv.py:61-128 with pc/RotationFunctions.py:16-21 (pair
                     order a*F_in+b), :637-665 (u = d^T R_out), :549-600 + :236-252 (rows 0-1 of
                     R_out^T R_in)
  conv_forward       layers/PNEConvLayerRotEquiv.py:199-216; aggregation semantics of
                     custom_ops/feature_aggregation/feat_basis_proj.cu:55-118 written as the scatter
                     formulation  T[r,c,k] = sum_e x[src(e),c] h[e,k]
  pca_frames         pc/RotationFunctions.py:307-406
  quat_frames        pc/RotationFunctions.py:176-216, 53-82
Backward comes from torch.autograd over conv_forward (the reference's backward is autograd over the
same graph plus feat_basis_proj_grads.cu, which computes exactly these sums).

Pinning: tests/golden/layer_*.npz, frames_*.npz are produced by tests/golden/gen_layer_golden.py by
running the reference's own Python (get_rot_tenors, __compute_convolution__,
sample_reference_frames_pca, sample_reference_frames) in this container; tests/test_oracle.py
checks this file against them.
"""
import math

import torch


def act_fn(name):
    return {"mlp_linear": lambda t: t, "mlp_relu": torch.relu,
            "mlp_gelu": lambda t: 0.5 * t * (1.0 + torch.erf(t / math.sqrt(2.0))), "mlp_sin": torch.sin}[name]


def rot_tensors(pts_in, pts_out, frames_in, frames_out, neighbors, norm):
    """g [E,F_out,F_in,9], expanded rows [E,F_out,F_in] (= i*F_out+a), cols (= j*F_in+b)."""
    i, j = neighbors[:, 0].long(), neighbors[:, 1].long()
    e = i.shape[0]
    fo, fi = frames_out.shape[1], frames_in.shape[1]
    d = (pts_in[j] - pts_out[i]) * norm
    Ro = frames_out[i].reshape(e, fo, 3, 3)
    Ri = frames_in[j].reshape(e, fi, 3, 3)
    u = torch.einsum("er,earc->eac", d, Ro)
    rel = torch.einsum("earm,ebrn->eabmn", Ro, Ri)
    r6 = rel[:, :, :, :2, :].reshape(e, fo, fi, 6)
    g = torch.cat((u[:, :, None, :].expand(e, fo, fi, 3), r6), dim=-1)
    a = torch.arange(fo)[None, :, None]
    b = torch.arange(fi)[None, None, :]
    rows = (i[:, None, None] * fo + a).expand(e, fo, fi)
    cols = (j[:, None, None] * fi + b).expand(e, fo, fi)
    return g, rows, cols


def conv_forward(x, proj_axes, proj_biases, conv_weights, pts_in, pts_out, frames_in, frames_out, neighbors,
                 norm_neigh_dist, norm_num_neighs, pne_type="mlp_gelu", chunk=1 << 15):
    """y [M*F_out, Cout]; differentiable wrt x and the three parameters."""
    fo, fi = frames_out.shape[1], frames_in.shape[1]
    m = pts_out.shape[0]
    cin, k, cout = conv_weights.shape
    act = act_fn(pne_type)
    T = torch.zeros((m * fo, cin, k), dtype=x.dtype)
    for s in range(0, neighbors.shape[0], chunk):
        g, rows, cols = rot_tensors(pts_in, pts_out, frames_in, frames_out, neighbors[s:s + chunk], norm_neigh_dist)
        h = act(g.reshape(-1, 9) @ proj_axes + proj_biases)
        contrib = x[cols.reshape(-1)][:, :, None] * h[:, None, :]
        T = T.index_add(0, rows.reshape(-1), contrib)
    y = torch.einsum("nik,iko->no", T, conv_weights)
    return y / fi * norm_num_neighs


def pca_frames(points, knn, fixed_axis=None):
    """[N, 4 or 2, 9] candidate frames; knn [N,k] (negative -> self)."""
    n, k = knn.shape
    idx = torch.where(knn < 0, torch.arange(n)[:, None].expand(n, k), knn.long())
    X = points[idx].clone()
    fixed = bool(fixed_axis)
    if fixed:
        X[:, :, int(fixed_axis)] = 0
    Xc = X - X.mean(dim=1, keepdim=True)
    C = Xc.transpose(1, 2) @ Xc
    _, vec = torch.linalg.eigh(C)
    if fixed:
        vec = torch.flip(vec, dims=[-1])
    vec = vec * torch.where(torch.linalg.det(vec) < 0, -1.0, 1.0)[:, None, None].to(vec.dtype)
    signs = [(1, 1, 1), (-1, -1, 1)] if fixed else [(1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)]
    S = torch.tensor(signs, dtype=vec.dtype)
    fr = vec[:, None, :, :] * S[None, :, None, :]
    if fixed and int(fixed_axis) == 1:
        fr = fr[:, :, :, [0, 2, 1]]
    if fixed:
        fr = torch.where(fr.abs() < 1e-6, torch.zeros_like(fr), fr)
    return fr.reshape(n, len(signs), 9)


def quat_frames(o):
    """[n,4] normal samples -> [n,9] rotation matrices."""
    s = (o * o).sum(1)
    nrm = torch.sqrt(s)
    nrm = torch.where((nrm < 0) != (o[:, 0] < 0), -nrm, nrm)
    q = o / nrm[:, None]
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    return torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)


def frame_set_distance(a, b):
    """Max over points of the set distance between two frame sets [N,F,9] (order-free)."""
    d = (a[:, :, None, :] - b[:, None, :, :]).abs().amax(-1)   # [N,F,F]
    return torch.maximum(d.amin(2).amax(1), d.amin(1).amax(1))


def standard_conv_forward(x, proj_axes, proj_biases, conv_weights, pts_in, pts_out, neighbors, norm_neigh_dist,
                          norm_num_neighs, pne_type="mlp_gelu"):
    """The non-equivariant PNEConvLayer with "add" aggregation: LinearPNE
    (custom_ops/PNE.py:30-40: rel = (p_in[j] - p_out[i]) * norm; basis = rel @ A + b), activation, the scatter
    formulation of FeatBasisProj, einsum with conv_weights_ and the norm_num_neighs_ scale
    (layers/PNEConvLayer.py:161-229).  y [M, Cout]; differentiable wrt x and the three parameters."""
    i, j = neighbors[:, 0].long(), neighbors[:, 1].long()
    rel = (pts_in[j] - pts_out[i]) * norm_neigh_dist
    h = act_fn(pne_type)(rel @ proj_axes + proj_biases[None, :])
    cin, k, cout = conv_weights.shape
    T = torch.zeros((pts_out.shape[0], cin, k), dtype=x.dtype)
    T.index_add_(0, i, x[j][:, :, None] * h[:, None, :])
    return torch.einsum("nik,iko->no", T, conv_weights) * norm_num_neighs


def act_grad_fn(name):
    return {"mlp_linear": lambda t: torch.ones_like(t), "mlp_relu": lambda t: (t > 0).to(t.dtype),
            "mlp_gelu": lambda t: 0.5 * (1.0 + torch.erf(t / math.sqrt(2.0))) +
            t * torch.exp(-0.5 * t * t) / math.sqrt(2.0 * math.pi),
            "mlp_sin": torch.cos}[name]


def conv_forward_backward(x, proj_axes, proj_biases, conv_weights, pts_in, pts_out, frames_in, frames_out, neighbors,
                          norm_neigh_dist, norm_num_neighs, dy, pne_type="mlp_gelu", chunk=1 << 13):
    """y and the four gradients (dx, dW, dA, dB) of sum(y * dy), WITHOUT autograd: the same sums as conv_forward and
    its autograd graph (layers/PNEConvLayerRotEquiv.py:199-216; backward of FeatBasisProj,
    custom_ops/feature_aggregation/feat_basis_proj_grads.cu:100-145: dBasis[e,k] = sum_c dT[r,c,k] x[src,c],
    dX[src,c] = sum_k dT[r,c,k] basis[e,k]), evaluated chunk by chunk so that BASELINE-size layers
    (E*C of a few million pairs) fit in host memory.  tests/test_oracle.py pins it to conv_forward + autograd."""
    fo, fi = frames_out.shape[1], frames_in.shape[1]
    m = pts_out.shape[0]
    cin, k, cout = conv_weights.shape
    act, dact = act_fn(pne_type), act_grad_fn(pne_type)
    s = norm_num_neighs / fi
    T = torch.zeros((m * fo, cin, k), dtype=x.dtype)
    for c0 in range(0, neighbors.shape[0], chunk):
        g, rows, cols = rot_tensors(pts_in, pts_out, frames_in, frames_out, neighbors[c0:c0 + chunk], norm_neigh_dist)
        h = act(g.reshape(-1, 9) @ proj_axes + proj_biases)
        T.index_add_(0, rows.reshape(-1), x[cols.reshape(-1)][:, :, None] * h[:, None, :])
    W2 = conv_weights.reshape(cin * k, cout)
    y = (T.reshape(m * fo, cin * k) @ W2) * s
    dW = (T.reshape(m * fo, cin * k).t() @ dy).reshape(cin, k, cout) * s
    dT = ((dy @ W2.t()) * s).reshape(m * fo, cin, k)
    del T
    dx = torch.zeros_like(x)
    dA = torch.zeros_like(proj_axes)
    dB = torch.zeros_like(proj_biases)
    for c0 in range(0, neighbors.shape[0], chunk):
        g, rows, cols = rot_tensors(pts_in, pts_out, frames_in, frames_out, neighbors[c0:c0 + chunk], norm_neigh_dist)
        g = g.reshape(-1, 9)
        rows, cols = rows.reshape(-1), cols.reshape(-1)
        pre = g @ proj_axes + proj_biases
        dTr = dT[rows]                                             # [n, cin, k]
        dx.index_add_(0, cols, torch.einsum("nck,nk->nc", dTr, act(pre)))
        dpre = torch.einsum("nck,nc->nk", dTr, x[cols]) * dact(pre)
        dA += g.t() @ dpre
        dB += dpre.sum(0)
    return y, dx, dW, dA, dB


def err_metrics(got, ref):
    """(max-abs error / max-abs value, ||err||_2 / ||ref||_2, 99.9th percentile of |err| / (|ref| + rms(ref)))
    -- the second and third cannot hide behind one large entry."""
    import numpy as np
    a, b = np.asarray(got, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    d = np.abs(a - b)
    rms = max(float(np.sqrt(np.mean(b * b))), 1e-300)
    return (float(d.max() / max(np.abs(b).max(), 1e-300)), float(np.sqrt((d * d).sum()) / max(np.sqrt((b * b).sum()), 1e-300)),
            float(np.percentile(d / (np.abs(b) + rms), 99.9)))
