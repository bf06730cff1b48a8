"""CPU tests pinning the integer oracle (oracle/se3_oracle.c) and the aggregation formulation to
outputs of the UNMODIFIED reference CUDA ops, frozen on a B200 by tests/golden/gen_ref_ops_golden.py."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import int_oracle as io

G = dict(np.load(os.path.join(GOLDEN, "ref_ops_golden.npz")))


def test_ball_query_and_keys_oracle_equals_reference_cuda():
    for name in ("bq_same", "bq_cross", "bq_flat"):
        src, dst, bs, bd = G[name + "_src"], G[name + "_dst"], G[name + "_bs"], G[name + "_bd"]
        r = float(G[name + "_r"])
        mn, nc = io.grid_setup_ball_query(src, bs, r)
        np.testing.assert_array_equal(mn, G[name + "_min"])
        np.testing.assert_array_equal(nc, G[name + "_nc"])
        rad = np.full(3, r, np.float32)
        nb, ends = io.ball_query(src, dst, bs, bd, mn, nc, rad)
        np.testing.assert_array_equal(ends, G[name + "_ends"])
        np.testing.assert_array_equal(io.canonical_rows(nb, ends), G[name + "_nb"])
        np.testing.assert_array_equal(io.compute_keys(src, bs, mn, nc, rad), G[name + "_keys"])


def test_knn_oracle_equals_reference_cuda():
    for name in ("knn_a", "knn_b"):
        pts, b, ref = G[name + "_pts"], G[name + "_b"], G[name + "_idx"]
        got, dist = io.knn_query(pts, b, ref.shape[1])
        np.testing.assert_array_equal(got < 0, ref < 0)
        P = pts.astype(np.float64)
        rd = np.where(ref >= 0, ((P[np.maximum(ref, 0)] - P[:, None, :]) ** 2).sum(-1), np.inf)
        gd = np.where(got >= 0, dist.astype(np.float64), np.inf)
        np.testing.assert_allclose(gd, rd, rtol=1e-5, atol=1e-12)   # same distances in the same slots
        assert (got != ref).mean() < 1e-3                            # indices equal up to exact ties


def test_scatter_formulation_equals_reference_feat_basis_proj():
    """The oracle's aggregation (scatter formulation) against the reference's CUDA op and its backward."""
    basis = torch.from_numpy(G["fbp_basis"]).double().requires_grad_(True)
    feats = torch.from_numpy(G["fbp_feats"]).double().requires_grad_(True)
    nbr = torch.from_numpy(G["fbp_nbr"]).long()
    m = G["fbp_ends"].shape[0]
    T = torch.zeros(m, feats.shape[1], basis.shape[1], dtype=torch.float64).index_add(
        0, nbr[:, 0], feats[nbr[:, 1]][:, :, None] * basis[:, None, :])
    (T * torch.from_numpy(G["fbp_grads"]).double()).sum().backward()
    np.testing.assert_allclose(T.detach().numpy(), G["fbp_T"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(feats.grad.numpy(), G["fbp_fg"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(basis.grad.numpy(), G["fbp_bg"], rtol=1e-4, atol=1e-4)
