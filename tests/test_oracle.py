"""CPU tests: the oracle (oracle/) against the golden vectors produced by the reference's own
Python (tests/golden/gen_layer_golden.py), plus self-consistency of the C integer oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, LAYER_CASES
from oracle import layer_oracle as lo
from oracle import int_oracle as io


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.mark.parametrize("case", LAYER_CASES)
def test_rot_tensors_match_reference(case):
    g = load("layer_%s.npz" % case)
    t = lambda k: torch.from_numpy(g[k]).double()
    geo, rows, cols = lo.rot_tensors(t("pts_in"), t("pts_out"), t("frames_in"), t("frames_out"),
                                     torch.from_numpy(g["neighbors"]), float(g["norm_neigh_dist"]))
    # the reference sorts the expanded list by row with an unstable sort: compare as multisets per (row, col)
    mine = np.concatenate([rows.reshape(-1, 1).numpy(), cols.reshape(-1, 1).numpy(), geo.reshape(-1, 9).numpy()], 1)
    ref = np.concatenate([g["nb_expanded"].astype(np.float64), g["g_sorted"]], 1)
    mine = mine[np.lexsort((mine[:, 1], mine[:, 0]))]
    ref = ref[np.lexsort((ref[:, 1], ref[:, 0]))]
    assert mine.shape == ref.shape
    np.testing.assert_array_equal(mine[:, :2], ref[:, :2])
    np.testing.assert_allclose(mine[:, 2:], ref[:, 2:], rtol=0, atol=1e-6)
    # inclusive end offsets of the expanded rows
    counts = np.bincount(rows.reshape(-1).numpy(), minlength=int(g["ends_expanded"].shape[0]))
    np.testing.assert_array_equal(np.cumsum(counts), g["ends_expanded"])


@pytest.mark.parametrize("case", LAYER_CASES)
def test_layer_forward_backward_match_reference(case):
    g = load("layer_%s.npz" % case)
    t = lambda k: torch.from_numpy(g[k]).double()
    x = t("x").requires_grad_(True)
    A, B, W = t("proj_axes").requires_grad_(True), t("proj_biases").requires_grad_(True), t("conv_weights").requires_grad_(True)
    y = lo.conv_forward(x, A, B, W, t("pts_in"), t("pts_out"), t("frames_in"), t("frames_out"),
                        torch.from_numpy(g["neighbors"]), float(g["norm_neigh_dist"]), float(g["norm_num_neighs"]),
                        str(g["pne"]))
    (y * t("dy")).sum().backward()
    tol = dict(rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(y.detach().numpy(), g["y_f64"], **tol)
    np.testing.assert_allclose(x.grad.numpy(), g["dx_f64"], **tol)
    np.testing.assert_allclose(W.grad.numpy(), g["dW_f64"], **tol)
    np.testing.assert_allclose(A.grad.numpy(), g["dA_f64"], **tol)
    np.testing.assert_allclose(B.grad.numpy(), g["dB_f64"], **tol)
    # the reference's own fp32 run agrees with its fp64 run to ~1e-5: that is the noise floor of "1e-4 rel"
    rel = np.abs(g["y_f32"] - g["y_f64"]).max() / np.abs(g["y_f64"]).max()
    assert rel < 1e-5


def test_pca_frames_match_reference():
    g = load("frames.npz")
    pts, knn = torch.from_numpy(g["pts"]).double(), torch.from_numpy(g["knn"])
    for tag, axis in (("none", None), ("axis2", 2), ("axis1", 1)):
        mine = lo.pca_frames(pts, knn, axis)
        ref = torch.from_numpy(g["frames64_" + tag])
        assert mine.shape == ref.shape
        # eigenvector signs are backend-defined: the frame SET is the invariant
        assert float(lo.frame_set_distance(mine, ref).max()) < 1e-6
        ref32 = torch.from_numpy(g["frames_" + tag]).double()
        assert float(lo.frame_set_distance(mine, ref32).max()) < 5e-3
        R = mine.reshape(-1, 3, 3)
        assert torch.allclose(R.transpose(1, 2) @ R, torch.eye(3, dtype=R.dtype).expand_as(R), atol=1e-6)
        # axis 1 applies the odd column permutation [0,2,1] (RotationFunctions.py:401-402): improper frames, as in the reference
        det = torch.linalg.det(R)
        assert torch.all((det if axis != 1 else -det) > 0.999)


def test_quat_frames_match_reference():
    g = load("frames.npz")
    mine = lo.quat_frames(torch.from_numpy(g["mc_randn"]))
    np.testing.assert_allclose(mine.numpy().reshape(50, 4, 9), g["mc_frames"], rtol=0, atol=1e-6)


def _cloud(n, b, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    pts = (rng.random((n, 3)) * scale).astype(np.float32)
    batch = np.sort(rng.integers(0, b, n)).astype(np.int32)
    return pts, batch


def test_int_oracle_ball_query_properties():
    src, bs = _cloud(700, 3, 0)
    dst, bd = _cloud(300, 3, 1)
    r = 0.17
    mn, nc = io.grid_setup_ball_query(src, bs, r)
    nb, ends = io.ball_query(src, dst, bs, bd, mn, nc, np.full(3, r, np.float32))
    assert ends[-1] == nb.shape[0] and np.all(np.diff(ends) >= 0)
    assert np.all(np.diff(nb[:, 0]) >= 0)
    # every pair is same-batch and inside the radius; pairs strictly inside the radius and inside the grid are all found
    d = np.linalg.norm(dst[nb[:, 0]].astype(np.float64) - src[nb[:, 1]].astype(np.float64), axis=1)
    assert np.all(bd[nb[:, 0]] == bs[nb[:, 1]]) and np.all(d < r * (1 + 1e-5))
    full = np.linalg.norm(dst[:, None, :].astype(np.float64) - src[None, :, :], axis=2)
    inside = np.all((dst >= mn[bd]) & (dst <= mn[bd] + nc * np.float32(r)), axis=1)
    expect = (full < r * (1 - 1e-5)) & (bd[:, None] == bs[None, :]) & inside[:, None]
    got = np.zeros_like(expect)
    got[nb[:, 0], nb[:, 1]] = True
    assert np.all(got[expect])
    # empty inputs
    nb0, ends0 = io.ball_query(src, dst[:0], bs, bd[:0], mn, nc, np.full(3, r, np.float32))
    assert nb0.shape == (0, 2) and ends0.shape == (0,)


def test_int_oracle_keys_and_knn():
    pts, b = _cloud(500, 2, 3)
    mn = np.stack([pts[b == i].min(0) for i in range(2)]) - np.float32(1e-6)
    nc = np.array([7, 7, 7], np.int32)
    keys = io.compute_keys(pts, b, mn, nc, np.full(3, 0.15, np.float32))
    cells = np.clip(np.floor((pts - mn[b]) * (np.float32(1.0) / np.float32(0.15))).astype(np.int64), 0, 6)
    np.testing.assert_array_equal(keys, ((b * 7 + cells[:, 0]) * 7 + cells[:, 1]) * 7 + cells[:, 2])
    k = 16
    idx, dist = io.knn_query(pts, b, k)
    full = ((pts[:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(2)
    full[b[:, None] != b[None, :]] = np.inf
    ref = np.sort(full, axis=1)[:, :k]
    np.testing.assert_allclose(dist, ref, rtol=1e-5, atol=1e-9)
    assert np.all(idx[:, 0] == np.arange(500))
    # a batch item smaller than k pads with -1
    small = np.concatenate([pts[:5], pts[100:]]).astype(np.float32)
    sb = np.concatenate([np.zeros(5, np.int32), np.ones(400, np.int32)])
    idx2, _ = io.knn_query(small, sb, k)
    assert np.all(idx2[:5, 5:] == -1) and np.all(idx2[:5, :5] >= 0)


@pytest.mark.parametrize("case", LAYER_CASES)
def test_chunked_forward_backward_matches_reference(case):
    """conv_forward_backward (the no-autograd, chunked restatement used at BASELINE sizes) against the reference's
    own float64 forward / backward, with a chunk size that splits the edge list."""
    g = load("layer_%s.npz" % case)
    t = lambda k: torch.from_numpy(g[k]).double()
    out = lo.conv_forward_backward(t("x"), t("proj_axes"), t("proj_biases"), t("conv_weights"), t("pts_in"), t("pts_out"),
                                   t("frames_in"), t("frames_out"), torch.from_numpy(g["neighbors"]),
                                   float(g["norm_neigh_dist"]), float(g["norm_num_neighs"]), t("dy"), str(g["pne"]),
                                   chunk=97)
    for got, key in zip(out, ("y", "dx", "dW", "dA", "dB")):
        np.testing.assert_allclose(got.numpy(), g[key + "_f64"], rtol=1e-9, atol=1e-10, err_msg=key)
        assert max(lo.err_metrics(got.numpy(), g[key + "_f64"])) < 1e-9
