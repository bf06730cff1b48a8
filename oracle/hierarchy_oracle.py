"""TEST INFRASTRUCTURE ONLY -- CPU port of the neighbourhood / reference-frame / hierarchy machinery
(numpy + oracle/se3_oracle.c + torch.linalg.eigh) and of one dfaust conv-stack step, used by
bench.py's cpu_baseline / `--impl reference` legs and by tests.  Never imported by se3conv3d_b200/.

Follows (paths relative to /root/reference/point_cloud_lib/point_cloud_lib): pc/Grid.py:26-58,
pc/BoundingBox.py:17-18, pc/GridSubSample.py:59-72 (grid average pooling), pc/KnnNeighborhood.py:39-75,
pc/PointcloudRotEquiv.py:77-178 (PCA frames; the random per-point frame permutation is replaced by
"first n_frames" so runs are reproducible -- BASELINE.md section 3), custom_ops/BallQuery.py:36-54.
"""
import numpy as np
import torch

from . import int_oracle as io
from . import layer_oracle as lo


def grid_pool(pts, batch, cell):
    b = int(batch.max()) + 1
    mn = np.stack([pts[batch == i].min(0) for i in range(b)]).astype(np.float32) - np.float32(1e-6)
    mx = np.stack([pts[batch == i].max(0) for i in range(b)]).astype(np.float32) + np.float32(1e-6)
    nc = (((mx - mn) / np.float32(cell)).astype(np.int32) + 1).max(0).astype(np.int32)
    keys = io.compute_keys(pts, batch, mn, nc, np.full(3, cell, np.float32))
    uniq, inv = np.unique(keys, return_inverse=True)
    cnt = np.bincount(inv, minlength=len(uniq)).astype(np.float32)
    pooled = np.stack([np.bincount(inv, weights=pts[:, d], minlength=len(uniq)) for d in range(3)], 1)
    pooled = (pooled / cnt[:, None]).astype(np.float32)
    pb = np.zeros(len(uniq), np.int32)
    np.maximum.at(pb, inv, batch.astype(np.int32))
    return pooled, pb, inv


class CpuCloud(object):
    def __init__(self, pts, batch, n_frames=2, k=16, fixed_axis=None):
        self.pts = np.ascontiguousarray(pts, np.float32)
        self.batch = np.ascontiguousarray(batch, np.int32)
        knn, _ = io.knn_query(self.pts, self.batch, k)
        cand = lo.pca_frames(torch.from_numpy(self.pts), torch.from_numpy(knn), fixed_axis)
        self.frames = cand[:, :n_frames, :].contiguous()
        self.n_frames = n_frames


def ball_query(src, dst, radius):
    mn, nc = io.grid_setup_ball_query(src.pts, src.batch, radius)
    nb, ends = io.ball_query(src.pts, dst.pts, src.batch, dst.batch, mn, nc, np.full(3, radius, np.float32))
    return torch.from_numpy(nb), ends


def dfaust_step_cpu(pts, batch, specs, params, cfg, n_frames=2):
    """Hierarchy + frames + neighbourhoods + the conv stack fwd+bwd on the host.  `params` is a list of
    (proj_axes, proj_biases, conv_weights) float32 tensors.  Returns (checksum, n_points)."""
    pts = np.ascontiguousarray(pts, np.float32)
    batch = np.ascontiguousarray(batch, np.int32)
    p0, b0, _ = grid_pool(pts, batch, cfg["init_subsample"])
    clouds = [CpuCloud(p0, b0, n_frames)]
    for cell in cfg["grid_subsamples"]:
        p, b, _ = grid_pool(clouds[-1].pts, clouds[-1].batch, cell)
        clouds.append(CpuCloud(p, b, n_frames))
    # output cloud: first point of every 0.04 cell (deterministic stand-in for the random pick)
    _, _, inv = grid_pool(pts, batch, cfg["init_subsample"])
    first = np.full(inv.max() + 1, -1, np.int64)
    first[inv[::-1]] = np.arange(len(inv))[::-1]
    clouds.append(CpuCloud(pts[first], batch[first], n_frames))
    radii = [cfg["init_subsample"]] + cfg["grid_subsamples"]
    cache = {}
    checksum = 0.0
    g = torch.Generator().manual_seed(1)
    for (name, li, lo_, lr, cin, cout), (A, B, W) in zip(specs, params):
        r = 2.0 * radii[lr]
        key = (li, lo_, r)
        if key not in cache:
            cache[key] = ball_query(clouds[li], clouds[lo_], r)
        nb, ends = cache[key]
        cin_pc, cout_pc = clouds[li], clouds[lo_]
        x = torch.randn(cin_pc.pts.shape[0] * n_frames, cin, generator=g).requires_grad_(True)
        dy = torch.randn(cout_pc.pts.shape[0] * n_frames, cout, generator=g)
        A, B, W = A.clone().requires_grad_(True), B.clone().requires_grad_(True), W.clone().requires_grad_(True)
        y = lo.conv_forward(x, A, B, W, torch.from_numpy(cin_pc.pts), torch.from_numpy(cout_pc.pts), cin_pc.frames,
                            cout_pc.frames, nb, 1.0 / r, len(ends) / max(nb.shape[0], 1), "mlp_gelu", chunk=4096)
        y.backward(dy)
        checksum += float(y.detach().sum())
    return checksum, pts.shape[0]
