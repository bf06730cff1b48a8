"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every product call goes through the
C ABI (libse3conv3d_b200.so); the checker is the oracle (oracle/), the golden vectors produced by
the reference's own Python, and -- when oracle/_ref was built -- the unmodified reference CUDA ops.

Tolerances:  integer / index results bit-exact;  fp32 mode outputs and gradients <= 1e-4 relative
(max-abs error over max-abs value);  bf16 tensor-core mode <= 3e-2 relative (stated with the test).
"""
import glob
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, LAYER_CASES, ROOT
from oracle import int_oracle as io
from oracle import layer_oracle as lo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def cloud(n, b, seed, scale=(1.0, 1.0, 1.0), shift=(0.0, 0.0, 0.0)):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g) * torch.tensor(scale) + torch.tensor(shift)
    batch = torch.sort(torch.randint(0, b, (n,), generator=g))[0].to(torch.int32)
    return pts, batch


def ref_ops():
    if not glob.glob(os.path.join(ROOT, "oracle", "_ref", "point_cloud_lib_ops*.so")):
        return None
    p = os.path.join(ROOT, "oracle", "_ref")
    if p not in sys.path:
        sys.path.insert(0, p)
    import point_cloud_lib_ops
    return point_cloud_lib_ops


def canon(nb):
    nb = np.asarray(nb)
    return nb[np.lexsort((nb[:, 1], nb[:, 0]))]


# ------------------------------------------------------------------------------------------------
def test_native_library_is_loaded():
    from se3conv3d_b200 import _lib
    assert _lib.lib().se3_abi_version() == 1
    maps = open("/proc/self/maps").read()
    assert "libse3conv3d_b200.so" in maps


@pytest.mark.parametrize("n,b,cell", [(5000, 1, 0.05), (7001, 4, 0.11), (3, 1, 0.5)])
def test_compute_keys_bit_exact(n, b, cell):
    from se3conv3d_b200 import point_cloud_lib_ops as ops
    pts, batch = cloud(n, b, 1)
    mn = torch.stack([pts[batch == i].min(0)[0] if (batch == i).any() else torch.zeros(3) for i in range(b)]) - 1e-6
    mx = torch.stack([pts[batch == i].max(0)[0] if (batch == i).any() else torch.zeros(3) for i in range(b)]) + 1e-6
    nc = torch.max(((mx - mn) / cell).to(torch.int32) + 1, dim=0)[0]
    cs = torch.full((3,), cell)
    got = ops.compute_keys(pts.to(DEV), batch.to(DEV), mn.to(DEV), nc.to(DEV), cs.to(DEV)).cpu().numpy()
    want = io.compute_keys(pts.numpy(), batch.numpy(), mn.numpy(), nc.numpy(), cs.numpy())
    np.testing.assert_array_equal(got, want)
    r = ref_ops()
    if r is not None:
        ref = r.compute_keys(pts.to(DEV), batch.to(DEV), mn.to(DEV), nc.to(DEV), cs.to(DEV)).cpu().numpy()
        np.testing.assert_array_equal(got, ref)


BQ_CASES = {
    "same": (lambda: cloud(4000, 3, 10), None, 0.1),
    "cross_outside_bbox": (lambda: cloud(3000, 2, 11), lambda: cloud(1500, 2, 12, (1.3, 1.3, 1.3), (-0.15,) * 3), 0.17),
    "flat": (lambda: cloud(2500, 1, 13, (1.0, 1.0, 0.02)), None, 0.08),
    "tiny": (lambda: cloud(5, 1, 14), None, 2.0),
}


@pytest.mark.parametrize("case", sorted(BQ_CASES))
def test_ball_query_bit_exact(case):
    from se3conv3d_b200.custom_ops import BallQuery
    mk_src, mk_dst, r = BQ_CASES[case]
    src, bs = mk_src()
    dst, bd = (src, bs) if mk_dst is None else mk_dst()
    nb, ends = BallQuery.apply(src.to(DEV), dst.to(DEV), bs.to(DEV), bd.to(DEV), r, 0)
    assert nb.dtype == torch.int64 and ends.dtype == torch.int32 and nb.shape[1] == 2
    nb, ends = nb.cpu().numpy(), ends.cpu().numpy()
    mn, nc = io.grid_setup_ball_query(src.numpy(), bs.numpy(), r)
    onb, oends = io.ball_query(src.numpy(), dst.numpy(), bs.numpy(), bd.numpy(), mn, nc, np.full(3, r, np.float32))
    np.testing.assert_array_equal(ends, oends)
    assert np.all(np.diff(nb[:, 0]) >= 0), "rows must be grouped by sample in increasing order"
    np.testing.assert_array_equal(canon(nb), canon(onb))
    # run-to-run determinism (the reference's order is an atomic race; ours is fixed)
    nb2, _ = BallQuery.apply(src.to(DEV), dst.to(DEV), bs.to(DEV), bd.to(DEV), r, 0)
    np.testing.assert_array_equal(nb, nb2.cpu().numpy())
    rops = ref_ops()
    if rops is not None:
        mnt = torch.from_numpy(mn).to(DEV)
        rnb, rends = rops.ball_query(src.to(DEV), dst.to(DEV), bs.to(DEV), bd.to(DEV), mnt,
                                     torch.from_numpy(nc).to(DEV), torch.full((3,), r, device=DEV), 0)
        np.testing.assert_array_equal(ends, rends.cpu().numpy())
        np.testing.assert_array_equal(canon(nb), canon(rnb.cpu().numpy()))


def test_ball_query_config1_edge_count():
    """BASELINE config 1: torch.rand(8192,3, seed 0), r = 0.1 -> E = 258,754 incl. self edges under the
    reference predicate length((s-p)*(1/r)) < 1 (oracle + reference CUDA op agree; SURVEY 8d's 258,756 came
    from torch.cdist, whose matmul-based distances flip two boundary pairs)."""
    from se3conv3d_b200.custom_ops import BallQuery
    pts = torch.rand(8192, 3, generator=torch.Generator().manual_seed(0))
    b = torch.zeros(8192, dtype=torch.int32)
    nb, ends = BallQuery.apply(pts.to(DEV), pts.to(DEV), b.to(DEV), b.to(DEV), 0.1, 0)
    assert nb.shape[0] == 258754 and int(ends[-1]) == 258754
    assert bool((nb[:, 0] == nb[:, 1]).sum() == 8192)


def test_ball_query_empty():
    from se3conv3d_b200.custom_ops import BallQuery
    src, bs = cloud(100, 1, 3)
    e3 = torch.zeros(0, 3)
    nb, ends = BallQuery.apply(src.to(DEV), e3.to(DEV), bs.to(DEV), torch.zeros(0, dtype=torch.int32, device=DEV), 0.1, 0)
    assert nb.shape == (0, 2) and ends.shape == (0,)


@pytest.mark.parametrize("n,b,k,scale", [(3000, 3, 16, (1.0, 0.5, 2.0)), (40, 4, 16, (1, 1, 1)), (20000, 32, 16, (0.6, 0.3, 1.8)),
                                         (500, 1, 32, (1, 1, 1)), (64, 1, 1, (1, 1, 1))])
def test_knn_matches_oracle(n, b, k, scale):
    from se3conv3d_b200.custom_ops import KNNQuery
    pts, batch = cloud(n, b, 20, scale)
    got = KNNQuery.apply(pts.to(DEV), batch.to(DEV), k).cpu().numpy()
    want, wdist = io.knn_query(pts.numpy(), batch.numpy(), k)
    assert got.dtype == np.int32 and got.shape == (n, k)
    np.testing.assert_array_equal(got < 0, want < 0)
    # identical indices except where equal distances tie (order of discovery among ties is not pinned)
    P = pts.numpy().astype(np.float64)
    gd = np.where(got >= 0, ((P[np.maximum(got, 0)] - P[:, None, :]) ** 2).sum(-1), np.inf)
    wd = np.where(want >= 0, wdist.astype(np.float64), np.inf)
    np.testing.assert_allclose(gd, wd, rtol=1e-5, atol=1e-12)
    mism = got != want
    assert mism.mean() < 1e-3
    r = ref_ops()
    if r is not None:
        ref = r.knn_query(pts.to(DEV), batch.to(DEV), k).cpu().numpy()
        rd = np.where(ref >= 0, ((P[np.maximum(ref, 0)] - P[:, None, :]) ** 2).sum(-1), np.inf)
        np.testing.assert_allclose(np.sort(gd, 1), np.sort(rd, 1), rtol=1e-5, atol=1e-12)


def test_pca_frames_match_oracle_and_golden():
    from se3conv3d_b200.pc import sample_reference_frames_pca
    g = dict(np.load(os.path.join(GOLDEN, "frames.npz")))
    pts = torch.from_numpy(g["pts"])
    knn = torch.from_numpy(g["knn"]).long()
    n, k = knn.shape
    nbr = torch.stack((torch.arange(n)[:, None].expand(n, k).reshape(-1), knn.reshape(-1)), 1)
    for tag, axis in (("none", False), ("axis2", 2), ("axis1", 1)):
        neigh = types.SimpleNamespace(neighbors_=nbr.to(DEV), k_=k)
        got = sample_reference_frames_pca(pts.to(DEV), neigh, axis_fixed=axis).cpu().double()
        ref = torch.from_numpy(g["frames64_" + tag])
        assert got.shape == ref.shape
        # Frames are compared as SETS up to the backend-defined eigenvector signs: on the 4-frame branch the set
        # is invariant; on the fixed-axis branch the sign of the (v0, fixed) column pair follows LAPACK's arbitrary
        # choice (RotationFunctions.py:374-383), so the admissible alternatives are the proper column-sign flips.
        # All eigengaps are healthy in this fixture.
        dist = None
        for sg in ((1, 1, 1), (-1, 1, -1), (1, -1, -1), (-1, -1, 1)):
            alt = (got.reshape(n, -1, 3, 3) * torch.tensor(sg, dtype=got.dtype)).reshape(got.shape)
            d = lo.frame_set_distance(alt, ref)
            dist = d if dist is None else torch.minimum(dist, d)
        assert float(dist.max()) < 2e-4
        R = got.reshape(-1, 3, 3)
        assert torch.allclose(R.transpose(1, 2) @ R, torch.eye(3, dtype=R.dtype).expand_as(R), atol=1e-5)


def test_mc_frames_match_golden():
    from se3conv3d_b200.pc import quaternion_to_matrix
    g = dict(np.load(os.path.join(GOLDEN, "frames.npz")))
    got = quaternion_to_matrix(torch.from_numpy(g["mc_randn"]).to(DEV)).cpu().numpy().reshape(50, 4, 9)
    np.testing.assert_allclose(got, g["mc_frames"], rtol=0, atol=2e-6)


def test_csr_transpose():
    from se3conv3d_b200.pc.neighborhood import ConvGeometry
    g = torch.Generator().manual_seed(5)
    m, n = 300, 200
    deg = torch.randint(0, 9, (m,), generator=g)
    rows = torch.repeat_interleave(torch.arange(m), deg)
    cols = torch.randint(0, n, (rows.shape[0],), generator=g)
    nb = torch.stack((rows, cols), 1)
    pc_in = types.SimpleNamespace(pts_=torch.rand(n, 3).to(DEV), local_frames_=torch.rand(n, 1, 9).to(DEV), n_frames_=1)
    pc_out = types.SimpleNamespace(pts_=torch.rand(m, 3).to(DEV), local_frames_=torch.rand(m, 1, 9).to(DEV), n_frames_=1)
    neigh = types.SimpleNamespace(neighbors_=nb.to(DEV), start_ids_=torch.cumsum(deg, 0).to(torch.int32).to(DEV))
    geo = ConvGeometry(pc_in, pc_out, neigh)
    np.testing.assert_array_equal(geo.col_src.cpu().numpy(), cols.numpy())
    order = np.lexsort((np.arange(rows.shape[0]), cols.numpy()))      # by source, stable in edge id
    np.testing.assert_array_equal(geo.t_edge.cpu().numpy(), order)
    np.testing.assert_array_equal(geo.t_dst.cpu().numpy(), rows.numpy()[order])
    np.testing.assert_array_equal(geo.t_row_ends.cpu().numpy(), np.cumsum(np.bincount(cols.numpy(), minlength=n)))


def _run_layer(g, precision, dev=DEV):
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    same = bool(g["same"])
    fi, fo = g["frames_in"].shape[1], g["frames_out"].shape[1]
    pc_in = types.SimpleNamespace(pts_=t("pts_in"), local_frames_=t("frames_in"), n_frames_=fi)
    pc_out = pc_in if same else types.SimpleNamespace(pts_=t("pts_out"), local_frames_=t("frames_out"), n_frames_=fo)
    from se3conv3d_b200.pc import BQNeighborhood
    neigh = BQNeighborhood.__new__(BQNeighborhood)
    neigh.neighbors_, neigh.start_ids_, neigh.conv_geometry_cache_ = t("neighbors"), t("ends"), {}
    cin, k, cout = g["conv_weights"].shape
    layer = PNEConvLayerRotEquiv(9, cin, cout, k, str(g["pne"])).to(dev)
    with torch.no_grad():
        layer.proj_axes_.copy_(t("proj_axes"))
        layer.proj_biases_.copy_(t("proj_biases"))
        layer.conv_weights_.copy_(t("conv_weights"))
        layer.norm_neigh_dist_.fill_(float(g["norm_neigh_dist"]))
        layer.norm_num_neighs_.fill_(float(g["norm_num_neighs"]))
    layer.precision = precision
    x = t("x").clone().requires_grad_(True)
    y = layer(pc_in, pc_out, x, neigh)
    (y * t("dy")).sum().backward()
    return (y.detach().cpu().numpy(), x.grad.cpu().numpy(), layer.conv_weights_.grad.cpu().numpy(),
            layer.proj_axes_.grad.cpu().numpy(), layer.proj_biases_.grad.cpu().numpy())


@pytest.mark.parametrize("case", LAYER_CASES)
def test_layer_fp32_matches_reference_golden(case):
    """fp32 exactness mode vs the reference's own forward/backward (fp64 golden): <= 1e-4 relative."""
    g = dict(np.load(os.path.join(GOLDEN, "layer_%s.npz" % case)))
    y, dx, dW, dA, dB = _run_layer(g, 0)
    for got, key in ((y, "y"), (dx, "dx"), (dW, "dW"), (dA, "dA"), (dB, "dB")):
        assert got.shape == g[key + "_f64"].shape
        assert rel_err(got, g[key + "_f64"]) < 1e-4, key


@pytest.mark.parametrize("case", LAYER_CASES)
def test_layer_bf16_matches_reference_golden(case):
    """bf16 tensor-core mode (bf16 operands, fp32 accumulation) vs the reference golden: stated tolerance
    3e-2 relative (max-abs error / max-abs value) on outputs and every gradient."""
    g = dict(np.load(os.path.join(GOLDEN, "layer_%s.npz" % case)))
    y, dx, dW, dA, dB = _run_layer(g, 1)
    for got, key in ((y, "y"), (dx, "dx"), (dW, "dW"), (dA, "dA"), (dB, "dB")):
        assert got.shape == g[key + "_f64"].shape
        err = rel_err(got, g[key + "_f64"])
        print(case, key, "bf16 rel err %.2e" % err)
        assert err < 3e-2, key


def test_layer_bf16_matches_fp32_path_config1():
    """BASELINE config 1 size: the tensor-core path against the fp32 path on the GPU (same inputs)."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    pc, neigh, x = _synthetic_layer_problem(8192, 0.1, 2, 32, 64)
    torch.manual_seed(2)
    layer = PNEConvLayerRotEquiv(9, 32, 64, 32, "mlp_gelu").to(DEV)
    layer.norm_neigh_dist_.fill_(10.0)
    layer.norm_num_neighs_.fill_(8192 / 258754)
    res = []
    for precision in (0, 1):
        layer.precision = precision
        layer.zero_grad()
        xx = x.clone().requires_grad_(True)
        y = layer(pc, pc, xx, neigh)
        y.square().mean().backward()
        res.append([t.detach().cpu().numpy() for t in (y, xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad,
                                                       layer.proj_biases_.grad)])
    for a, b, name in zip(res[1], res[0], ("y", "dx", "dW", "dA", "dB")):
        err = rel_err(a, b)
        print(name, "bf16 vs fp32 rel err %.2e" % err)
        assert err < 3e-2, name


def _synthetic_layer_problem(n, r, fi, cin, cout, seed=0, batches=1):
    from se3conv3d_b200.pc import PointcloudRotEquiv, BQNeighborhood
    torch.manual_seed(seed)
    pts = torch.rand(n, 3, generator=torch.Generator().manual_seed(seed))
    batch = torch.sort(torch.randint(0, batches, (n,), generator=torch.Generator().manual_seed(seed + 1)))[0].to(torch.int32)
    cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": fi}
    pc = PointcloudRotEquiv(pts.to(DEV), batch.to(DEV), cfg)
    neigh = BQNeighborhood(pc, pc, r)
    x = torch.randn(n * fi, cin, generator=torch.Generator().manual_seed(seed + 2)).to(DEV)
    return pc, neigh, x


def test_layer_fp32_matches_oracle_medium():
    """2048 points, F=2, 32->64: CUDA fp32 path vs the CPU oracle in float64 on identical inputs."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    pc, neigh, x = _synthetic_layer_problem(2048, 0.16, 2, 32, 64)
    torch.manual_seed(2)
    layer = PNEConvLayerRotEquiv(9, 32, 64, 32, "mlp_gelu").to(DEV)
    layer.norm_neigh_dist_.fill_(1 / 0.16)
    layer.norm_num_neighs_.fill_(2048 / neigh.neighbors_.shape[0])
    x = x.requires_grad_(True)
    y = layer(pc, pc, x, neigh)
    loss = y.square().mean()
    loss.backward()
    c = lambda t: t.detach().cpu().double()
    xo = c(x).requires_grad_(True)
    A, B, W = c(layer.proj_axes_).requires_grad_(True), c(layer.proj_biases_).requires_grad_(True), c(layer.conv_weights_).requires_grad_(True)
    yo = lo.conv_forward(xo, A, B, W, c(pc.pts_), c(pc.pts_), c(pc.local_frames_), c(pc.local_frames_),
                         neigh.neighbors_.cpu(), float(layer.norm_neigh_dist_), float(layer.norm_num_neighs_))
    yo.square().mean().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < 1e-4
    assert rel_err(x.grad.cpu().numpy(), xo.grad.numpy()) < 1e-4
    assert rel_err(layer.conv_weights_.grad.cpu().numpy(), W.grad.numpy()) < 1e-4
    assert rel_err(layer.proj_axes_.grad.cpu().numpy(), A.grad.numpy()) < 1e-4
    assert rel_err(layer.proj_biases_.grad.cpu().numpy(), B.grad.numpy()) < 1e-4


def test_layer_full_size_properties():
    """BASELINE config 1 size (8192 points, F=2, 32->64): size-independent properties --
    linearity in x, determinism, and frame-pooled equivariance under a global rotation."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    from se3conv3d_b200.pc import random_rotation
    pc, neigh, x = _synthetic_layer_problem(8192, 0.1, 2, 32, 64)
    assert neigh.neighbors_.shape[0] == 258754
    torch.manual_seed(2)
    layer = PNEConvLayerRotEquiv(9, 32, 64, 32, "mlp_gelu").to(DEV)
    layer.norm_neigh_dist_.fill_(10.0)
    layer.norm_num_neighs_.fill_(8192 / 258754)
    with torch.no_grad():
        y1 = layer(pc, pc, x, neigh)
        x2 = torch.randn_like(x)
        y2 = layer(pc, pc, x2, neigh)
        y12 = layer(pc, pc, 0.5 * x - 2.0 * x2, neigh)
        assert rel_err((0.5 * y1 - 2.0 * y2).cpu().numpy(), y12.cpu().numpy()) < 1e-5
        assert torch.equal(y1, layer(pc, pc, x, neigh))
        # rotate the cloud and its frames: per-frame outputs are invariant (frames co-rotate)
        R = random_rotation(device=DEV)
        rot = types.SimpleNamespace(pts_=pc.pts_ @ R.T, n_frames_=2,
                                    local_frames_=torch.matmul(R, pc.local_frames_.reshape(-1, 2, 3, 3)).reshape(-1, 2, 9).contiguous())
        neigh.conv_geometry_cache_ = {}
        yr = layer(rot, rot, x, neigh)
        err = rel_err(pc.feature_pooling(yr).cpu().numpy(), pc.feature_pooling(y1).cpu().numpy())
        print("frame-pooled equivariance error:", err)
        assert err < 1e-4


def test_legacy_feat_basis_proj_ops():
    from se3conv3d_b200.custom_ops import FeatBasisProj
    g = torch.Generator().manual_seed(30)
    m, n, c, k = 300, 400, 16, 32
    deg = torch.randint(0, 12, (m,), generator=g)
    rows = torch.repeat_interleave(torch.arange(m), deg)
    cols = torch.randint(0, n, (rows.shape[0],), generator=g)
    nbr = torch.stack((rows, cols), 1)
    ends = torch.cumsum(deg, 0).to(torch.int32)
    basis = torch.randn(rows.shape[0], k, generator=g).to(DEV).requires_grad_(True)
    feats = torch.randn(n, c, generator=g).to(DEV).requires_grad_(True)
    grads = torch.randn(m, c, k, generator=g).to(DEV)
    T = FeatBasisProj.apply(basis, feats, nbr.to(DEV), ends.to(DEV))
    (T * grads).sum().backward()
    b2 = basis.detach().clone().requires_grad_(True)
    f2 = feats.detach().clone().requires_grad_(True)
    T2 = torch.zeros(m, c, k, device=DEV).index_add(0, rows.to(DEV), f2[cols.to(DEV)][:, :, None] * b2[:, None, :])
    (T2 * grads).sum().backward()
    assert rel_err(T.detach().cpu().numpy(), T2.detach().cpu().numpy()) < 1e-5
    assert rel_err(basis.grad.cpu().numpy(), b2.grad.cpu().numpy()) < 1e-5
    assert rel_err(feats.grad.cpu().numpy(), f2.grad.cpu().numpy()) < 1e-5
    r = ref_ops()
    if r is not None:
        Tr = r.feat_basis_proj(b2.detach(), f2.detach(), nbr.to(torch.int32).to(DEV), ends.to(DEV))
        assert rel_err(T.detach().cpu().numpy(), Tr.cpu().numpy()) < 1e-5


def test_grid_subsample_and_hierarchy():
    from se3conv3d_b200.pc import PointcloudRotEquiv, PointHierarchyRotEquiv, GridSubSample, Pointcloud
    pts, batch = cloud(6000, 4, 40, (0.6, 0.3, 1.8))
    pc0 = Pointcloud(pts.to(DEV), batch.to(DEV))
    samp = GridSubSample(pc0, 0.04)
    # cell ids are dense ranks of the oracle keys
    mn = torch.stack([pts[batch == i].min(0)[0] for i in range(4)]) - 1e-6
    mx = torch.stack([pts[batch == i].max(0)[0] for i in range(4)]) + 1e-6
    nc = torch.max(((mx - mn) / 0.04).to(torch.int32) + 1, dim=0)[0]
    keys = io.compute_keys(pts.numpy(), batch.numpy(), mn.numpy(), nc.numpy(), np.full(3, 0.04, np.float32))
    uniq, inv = np.unique(keys, return_inverse=True)
    np.testing.assert_array_equal(samp.grid_.cell_ids_.cpu().numpy(), inv)
    pooled = samp.__subsample_tensor__(pc0.pts_, "avg").cpu().numpy()
    want = np.stack([pts.numpy()[inv == c].astype(np.float64).mean(0) for c in range(len(uniq))])
    np.testing.assert_allclose(pooled, want, rtol=0, atol=1e-6)
    pb = samp.__subsample_tensor__(pc0.batch_ids_, "max").cpu().numpy()
    np.testing.assert_array_equal(pb, np.array([batch.numpy()[inv == c].max() for c in range(len(uniq))]))
    up = samp.__upsample_tensor__(torch.from_numpy(want).float().to(DEV))
    assert up.shape == (6000, 3)
    cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": 2}
    pc = PointcloudRotEquiv(torch.from_numpy(pooled).to(DEV), torch.from_numpy(pb).to(DEV), cfg)
    assert pc.local_frames_.shape == (len(uniq), 2, 9) and pc.n_frames_ == 2
    h = PointHierarchyRotEquiv(pc, 3, "grid_avg", grid_radii=[0.08, 0.16, 0.32])
    sizes = [p.pts_.shape[0] for p in h.pcs_]
    assert sizes == sorted(sizes, reverse=True) and len(sizes) == 4
    nb = h.create_neighborhood(0, 1, "ball_query", bq_radius=0.16)
    assert nb is h.create_neighborhood(0, 1, "ball_query", bq_radius=0.16)
    assert nb.start_ids_.shape[0] == sizes[1]
    R = h.pcs_[2].local_frames_.reshape(-1, 3, 3)
    assert torch.allclose(R.transpose(1, 2) @ R, torch.eye(3, device=DEV).expand_as(R), atol=1e-5)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", [0, 1, 2, 3])
@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (300, 32, 1024), (1000, 16, 8), (257, 256, 2048), (5000, 48, 520),
                                   (4096, 1024, 64)])
def test_projection_gemm_matches_fp64(impl, m, n, k):
    """The [K*Cin] x Cout projection on its own (auto = impl 0, mma.sync = 1, tcgen05 with cp.async loads = 2, the
    persistent TMA-fed tcgen05 kernel = 3): bf16 operands, fp32 accumulation; tolerance 2e-5 relative for fp32 output
    (accumulation order only), 1.5e-2 for bf16 output."""
    from se3conv3d_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device=DEV).manual_seed(m + n + k)
    a = torch.randn(m, k, device=DEV, generator=g).to(torch.bfloat16)
    b = torch.randn(n, k, device=DEV, generator=g).to(torch.bfloat16)
    ref = 0.25 * (a.double() @ b.double().t())
    for out_bf16, tol in ((0, 2e-5), (1, 1.5e-2)):
        c = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16 if out_bf16 else torch.float32)
        _lib.check(L.se3_gemm_bf16_tn(_lib.ptr(a), _lib.ptr(b), m, n, k, 0.25, _lib.ptr(c), out_bf16, impl,
                                      _lib.stream()), "se3_gemm_bf16_tn")
        torch.cuda.synchronize()
        assert rel_err(c.double().cpu(), ref.cpu()) < tol


@pytest.mark.parametrize("impl", [0, 1, 2, 3])
@pytest.mark.parametrize("m,n,k", [(1024, 32, 84468), (128, 64, 64), (1024, 72, 1), (2048, 64, 63), (512, 512, 3000),
                                   (8, 8, 100), (16384, 256, 777), (1000, 40, 5001)])
def test_weight_gradient_gemm_matches_fp64(impl, m, n, k):
    """dW = T^T dy on its own (se3_gemm_bf16_mn: operands stored [K][M] / [K][N], the contraction index is the row index):
    auto, mma.sync, the cp.async tcgen05 kernel and the persistent TMA-fed one (MN-major SWIZZLE_128B boxes) against
    fp64 -- seg_head's shape, single and partial k-blocks, N below one 64-column atom, partial row tiles, N = 512."""
    from se3conv3d_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device=DEV).manual_seed(m + n + k)
    a = torch.randn(k, m, device=DEV, generator=g).to(torch.bfloat16)
    b = torch.randn(k, n, device=DEV, generator=g).to(torch.bfloat16)
    ref = 0.5 * (a.double().t() @ b.double())
    c = torch.full((m, n), float("nan"), device=DEV, dtype=torch.float32)
    _lib.check(L.se3_gemm_bf16_mn(_lib.ptr(a), _lib.ptr(b), m, n, k, 0.5, _lib.ptr(c), impl, _lib.stream()), "se3_gemm_bf16_mn")
    torch.cuda.synchronize()
    err = rel_err(c.double().cpu(), ref.cpu())
    print("dW gemm", m, n, k, "impl", impl, "rel err %.2e" % err)
    # fp32 accumulation of k terms in one chain (the layer splits long k over CTAs): the bound grows with sqrt(k)
    assert err < 2e-5 * max(1.0, (k / 2048.0) ** 0.5)


@pytest.mark.parametrize("n_clouds,n_points,n_batches", [(4, 3000, 4), (2, 6890, 2), (3, 900, 5), (2, 8000, 2)])
def test_fused_hierarchy_matches_per_object_path(n_clouds, n_points, n_batches):
    """se3_hierarchy_build (one native call) against the per-object chain with the reference's API on the
    same synthetic bodies: level clouds, grids and every CSR bit-exact; frames drawn from the same PCA
    candidates; the output cloud picks one raw point of every init voxel.  The cases cover the per-item CTA sorts
    of every size class (6890 points: the 1024 x 7 configuration), trailing empty batch items (n_batches larger than
    the ids present) and clouds too large for a CTA (8000 points: the device-wide sorts for the raw grid)."""
    from se3conv3d_b200 import workloads as wl
    from se3conv3d_b200.pc import BQNeighborhood
    pts, b = wl.synthetic_bodies(n_clouds, n_points, seed=3)
    pts, b = pts.to(DEV), b.to(DEV)
    step = wl.DfaustStep(DEV, precision=1)
    pcs_u, neighs_u = step.build_hierarchy(pts, b, fused=False)
    pcs_f, neighs_f = step.build_hierarchy(pts, b, fused=True, n_batches=n_batches)
    h = step.hierarchy
    assert len(pcs_u) == len(pcs_f) == 6
    for lvl in range(5):
        assert torch.equal(pcs_u[lvl].pts_, pcs_f[lvl].pts_), lvl
        assert torch.equal(pcs_u[lvl].batch_ids_.to(torch.int32), pcs_f[lvl].batch_ids_), lvl
        cand = pcs_u[lvl].local_frames_pca_cache_["se3-all"]                       # [N,4,9]
        fr = pcs_f[lvl].local_frames_                                              # [N,2,9]
        match = (fr[:, :, None, :] == cand[:, None, :, :]).all(-1)                 # [N,2,4]
        assert bool(match.any(-1).all()), "fused frames must come from the PCA candidates"
        assert bool((match[:, 0].float().argmax(-1) != match[:, 1].float().argmax(-1)).all())
    # pooling / upsampling indices of every level
    hu = None
    for lvl in range(4):
        gu = wl_grid_of(pcs_u, lvl, step)
        gf = h.sub_sampled_objs_[lvl].grid_
        assert torch.equal(gu.cell_ids_, gf.cell_ids_) and torch.equal(gu.sorted_ids_, gf.sorted_ids_)
        assert torch.equal(gu.cell_ends_, gf.cell_ends_) and gu.num_used_cells_ == gf.num_used_cells_
    # output cloud: one raw point per init voxel, in voxel order
    out = pcs_f[5]
    assert out.pts_.shape[0] == pcs_f[0].pts_.shape[0]
    assert torch.equal(h.init_cell_ids_[out.picked_ids_], torch.arange(out.pts_.shape[0], device=DEV))
    assert torch.equal(out.pts_, pts[out.picked_ids_])
    # neighbourhoods on the fused clouds: CSR and transposed CSR equal the per-object ops
    for nb in {id(n): n for n in neighs_f}.values():
        ref = BQNeighborhood(nb.pc_src_, nb.samples_, nb.radius_)
        assert torch.equal(ref.start_ids_, nb.start_ids_)
        assert torch.equal(ref.neighbors_, nb.neighbors_)
        g_ref = ref.conv_geometry(nb.pc_src_, nb.samples_)
        g = list(nb.conv_geometry_cache_.values())[0]
        e = g.n_edges
        assert e == g_ref.n_edges
        for k in ("col_src", "t_edge", "t_dst"):
            assert torch.equal(getattr(g_ref, k)[:e], getattr(g, k)[:e]), k
        assert torch.equal(g_ref.t_row_ends[:g.n_in], g.t_row_ends[:g.n_in])
        assert torch.equal(g_ref.rec_in, g.rec_in) and torch.equal(g_ref.rec_out, g.rec_out)


def test_fused_hierarchy_long_transposed_rows():
    """Coarse sources against a fine cloud: transposed CSR rows tens of thousands of entries long (the row-ordering
    kernel's chunk sort + rank-merge path) and empty rows; bit-exact against the per-object transpose."""
    from se3conv3d_b200.pc import build_point_hierarchy, BQNeighborhood
    g = torch.Generator().manual_seed(41)
    n = 40000
    pts = torch.rand(n, 3, generator=g)
    batch = torch.sort(torch.randint(0, 2, (n,), generator=g))[0].to(torch.int32)
    cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": 2}
    wanted = [(1, 0, 1.0), (0, 1, 0.3), (1, 1, 0.6), (0, 0, 0.03)]
    h, _ = build_point_hierarchy(pts.to(DEV), batch.to(DEV), cfg, 0.02, [0.5], neighborhoods=wanted, n_batches=2)
    assert h.pcs_[1].pts_.shape[0] <= 16 and h.pcs_[0].pts_.shape[0] > 20000
    longest = 0
    for nb in h.fused_neighborhoods_:
        ref = BQNeighborhood(nb.pc_src_, nb.samples_, nb.radius_)
        g_ref = ref.conv_geometry(nb.pc_src_, nb.samples_)
        gg = list(nb.conv_geometry_cache_.values())[0]
        e = gg.n_edges
        assert e == g_ref.n_edges and torch.equal(ref.start_ids_, nb.start_ids_)
        for k in ("col_src", "t_edge", "t_dst"):
            assert torch.equal(getattr(g_ref, k)[:e], getattr(gg, k)[:e]), k
        assert torch.equal(g_ref.t_row_ends[:gg.n_in], gg.t_row_ends[:gg.n_in])
        te = torch.cat((torch.zeros(1, dtype=torch.int32, device=DEV), gg.t_row_ends[:gg.n_in]))
        longest = max(longest, int((te[1:] - te[:-1]).max()))
    assert longest > 4096, longest


def test_fused_hierarchy_objects_release_their_arena_without_gc():
    """The objects of a fused build hold no reference cycles: dropping them frees the arena by reference counting,
    so the next step reuses the block (a cycle would park ~300 MB per step until the cyclic collector runs)."""
    import gc
    import weakref
    from se3conv3d_b200 import workloads as wl
    pts, b = wl.synthetic_bodies(2, 2000, seed=5)
    step = wl.DfaustStep(DEV, precision=1)
    gc.collect()
    gc.disable()
    try:
        pcs, neighs = step.build_hierarchy(pts.to(DEV), b.to(DEV), fused=True, n_batches=2)
        _ = neighs[0].start_ids_, neighs[0].neighbors_, pcs[0]._se3_records        # touch a few lazy windows
        _ = list(neighs[3].conv_geometry_cache_.values())[0].t_dst, step.hierarchy.init_cell_ids_
        ref = weakref.ref(step.hierarchy.fused_arena_)
        del pcs, neighs, _
        step.hierarchy = None
        assert ref() is None, "the arena is still referenced after the hierarchy objects were dropped"
    finally:
        gc.enable()


def test_fused_hierarchy_against_cpu_oracle():
    """The fused builder against the C oracle directly (no GPU code on the reference side): the level-0 grid
    (oracle keys -> unique -> dense ranks, cell means in sorted order), the kNN table behind the PCA frames is not
    exposed, so frames are checked as candidates elsewhere; every ball-query CSR row by row."""
    from se3conv3d_b200 import workloads as wl
    pts, b = wl.synthetic_bodies(3, 1800, seed=7)
    step = wl.DfaustStep(DEV, precision=1)
    pcs, neighs = step.build_hierarchy(pts.to(DEV), b.to(DEV), fused=True, n_batches=3)
    h = step.hierarchy
    # level 0 = grid average of the raw cloud at 0.04 (pc/Grid.py:26-58, pc/GridSubSample.py:59-73)
    p_np, b_np = pts.numpy(), b.numpy()
    cell = np.float32(wl.DFAUST_CFG["init_subsample"])
    mn = np.stack([p_np[b_np == i].min(0) for i in range(3)]).astype(np.float32) - np.float32(1e-6)
    mx = np.stack([p_np[b_np == i].max(0) for i in range(3)]).astype(np.float32) + np.float32(1e-6)
    nc = (((mx - mn) * (np.float32(1.0) / cell)).astype(np.int32) + 1).max(0).astype(np.int32)
    keys = io.compute_keys(p_np, b_np, mn, nc, np.full(3, cell, np.float32))
    uniq, inv = np.unique(keys, return_inverse=True)
    assert np.array_equal(h.init_cell_ids_.cpu().numpy(), inv)
    assert pcs[0].pts_.shape[0] == uniq.shape[0]
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(h.init_sorted_ids_.cpu().numpy(), order)
    ends = np.cumsum(np.bincount(inv, minlength=uniq.shape[0]))
    assert np.array_equal(h.init_cell_ends_.cpu().numpy(), ends)
    means = np.zeros((uniq.shape[0], 3), np.float32)
    lo_ = 0
    for c, hi_ in enumerate(ends):                     # sequential fp32 sums in sorted order, then one division
        acc = np.zeros(3, np.float32)
        for i in order[lo_:hi_]:
            acc = (acc + p_np[i]).astype(np.float32)
        means[c] = acc / np.float32(hi_ - lo_)
        lo_ = hi_
    assert np.array_equal(pcs[0].pts_.cpu().numpy(), means)
    assert np.array_equal(pcs[0].batch_ids_.cpu().numpy(), b_np[order[ends - 1]])
    # every neighbourhood against the oracle ball query on the same clouds
    for nb in {id(n): n for n in neighs}.values():
        s_np, d_np = nb.pc_src_.pts_.cpu().numpy(), nb.samples_.pts_.cpu().numpy()
        bs, bd = nb.pc_src_.batch_ids_.cpu().numpy(), nb.samples_.batch_ids_.cpu().numpy()
        mn_q, nc_q = io.grid_setup_ball_query(s_np, bs, nb.radius_)
        ref_nb, ref_ends = io.ball_query(s_np, d_np, bs, bd, mn_q, nc_q, np.full(3, nb.radius_, np.float32))
        assert np.array_equal(nb.start_ids_.cpu().numpy(), ref_ends)
        assert np.array_equal(io.canonical_rows(nb.neighbors_.cpu().numpy(), None), io.canonical_rows(ref_nb, None))


def test_full_size_hierarchy_invariants():
    """BASELINE config 2 at full size (32 x 6890 points): properties that do not need a reference -- the dense cell
    ranks are consistent with the sorted order, pooled sizes chain, every CSR is grouped by sample with sources of
    the same batch item inside the radius, same-level neighbourhoods are symmetric, and every transposed CSR is the
    permutation of its forward CSR with rows in ascending edge order."""
    from se3conv3d_b200 import workloads as wl
    pts, b = wl.synthetic_bodies(32, 6890, seed=0)
    step = wl.DfaustStep(DEV, precision=1)
    pcs, neighs = step.build_hierarchy(pts.to(DEV), b.to(DEV), fused=True, n_batches=32)
    h = step.hierarchy
    sizes = [int(pc.pts_.shape[0]) for pc in pcs]
    assert sizes[0] == sizes[5] and all(a > c for a, c in zip(sizes[:5], sizes[1:5]))
    for lvl, samp in enumerate(h.sub_sampled_objs_):
        g = samp.grid_
        n, m = sizes[lvl], sizes[lvl + 1]
        ids, srt, ends = g.cell_ids_, g.sorted_ids_, g.cell_ends_.to(torch.int64)
        assert int(ids.max()) == m - 1 and int(ends[-1]) == n and bool((ends[1:] > ends[:-1]).all())
        assert torch.equal(torch.sort(srt)[0], torch.arange(n, device=DEV))          # a permutation
        ranks_sorted = ids[srt]
        assert bool((ranks_sorted[1:] >= ranks_sorted[:-1]).all())                    # ranks ascend in sorted order
        assert torch.equal(torch.bincount(ids, minlength=m).cumsum(0), ends)
        assert bool((pcs[lvl + 1].batch_ids_[1:] >= pcs[lvl + 1].batch_ids_[:-1]).all())
    for nb in {id(n): n for n in neighs}.values():
        geom = list(nb.conv_geometry_cache_.values())[0]
        e = geom.n_edges
        pairs, ends = nb.neighbors_, nb.start_ids_.to(torch.int64)
        assert int(ends[-1]) == e == pairs.shape[0] and bool((ends[1:] >= ends[:-1]).all())
        assert bool((pairs[1:, 0] >= pairs[:-1, 0]).all())
        src, dst = nb.pc_src_, nb.samples_
        assert torch.equal(src.batch_ids_[pairs[:, 1]], dst.batch_ids_[pairs[:, 0]])
        d = (src.pts_[pairs[:, 1]] - dst.pts_[pairs[:, 0]]).norm(dim=1)
        assert float(d.max()) < nb.radius_ * (1 + 1e-5)
        if src is dst:                                                                # symmetric relation
            n = src.pts_.shape[0]
            fwd = pairs[:, 0] * n + pairs[:, 1]
            rev = pairs[:, 1] * n + pairs[:, 0]
            assert torch.equal(torch.sort(fwd)[0], torch.sort(rev)[0])
        t_edge, t_dst = geom.t_edge[:e].to(torch.int64), geom.t_dst[:e].to(torch.int64)
        t_ends = geom.t_row_ends[:geom.n_in].to(torch.int64)
        assert torch.equal(torch.sort(t_edge)[0], torch.arange(e, device=DEV))         # a permutation of the edges
        assert torch.equal(t_dst, pairs[t_edge, 0]) and int(t_ends[-1]) == e
        rows = torch.repeat_interleave(torch.arange(geom.n_in, device=DEV), torch.diff(t_ends, prepend=t_ends.new_zeros(1)))
        assert torch.equal(rows, pairs[t_edge, 1])                                     # entry t belongs to row src(edge)
        same_row = rows[1:] == rows[:-1]
        assert bool((t_edge[1:][same_row] > t_edge[:-1][same_row]).all())              # ascending edge order per row


def test_conv_stack_full_size_deterministic_and_bf16_vs_fp32():
    """BASELINE config 2 at full size: the 21 convolutions forward + backward twice on the same hierarchy give
    bit-identical outputs and gradients (atomic-free backward, ordered reductions, dependent launches that wait for
    their predecessor); the largest layer (seg_head, E ~ 775 k) agrees between the bf16 and the fp32 path."""
    from se3conv3d_b200 import workloads as wl
    pts, b = wl.synthetic_bodies(32, 6890, seed=0)
    step = wl.DfaustStep(DEV, precision=1)
    pcs, neighs = step.build_hierarchy(pts.to(DEV), b.to(DEV), fused=True, n_batches=32)
    step.calibrate(pcs, neighs)
    xs, dys = step.make_inputs(pcs)
    runs = []
    for _ in range(2):
        step.zero_grad()
        ys = [layer(pcs[li], pcs[lo], x, nb)
              for layer, nb, (_, li, lo, _, _, _), x in zip(step.layers, neighs, step.specs, xs)]
        torch.autograd.backward(ys, list(dys))
        runs.append([y.detach().clone() for y in ys] + [x.grad.clone() for x in xs] +
                    [p.grad.clone() for layer in step.layers for p in layer.parameters()])
    for a, c in zip(*runs):
        assert torch.equal(a, c)
    i = [sp[0] for sp in step.specs].index("seg_head")
    layer, nb, (_, li, lo, _, _, _) = step.layers[i], neighs[i], step.specs[i]
    res = []
    for precision in (1, 0):
        layer.precision = precision
        layer.zero_grad()
        xx = xs[i].detach().clone().requires_grad_(True)
        y = layer(pcs[li], pcs[lo], xx, nb)
        y.backward(dys[i])
        res.append([t.detach().cpu().numpy() for t in (y, xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad,
                                                       layer.proj_biases_.grad)])
    for a, c, name in zip(res[0], res[1], ("y", "dx", "dW", "dA", "dB")):
        err = rel_err(a, c)
        print("seg_head full size", name, "bf16 vs fp32 rel err %.2e" % err)
        assert err < 3e-2, name


def wl_grid_of(pcs, lvl, step):
    """Grid of level `lvl` -> `lvl + 1` rebuilt with the per-object API (for comparison)."""
    from se3conv3d_b200 import workloads as wl
    from se3conv3d_b200.pc import Grid
    return Grid(pcs[lvl], wl.DFAUST_CFG["grid_subsamples"][lvl])


@pytest.mark.parametrize("cin,cout,frames", [(64, 64, 2), (128, 32, 2), (1, 32, 2), (16, 48, 1), (48, 64, 1), (256, 128, 2),
                                              (512, 64, 1), (320, 256, 2), (40, 8, 2), (160, 72, 4), (128, 128, 3)])
def test_layer_bf16_matches_fp32_path_channel_sweep(cin, cout, frames):
    """Every channel-block configuration of the tensor-core kernels (16 / 32 / 64-channel row items, multi-block
    items, padded odd channel counts, the wide-layer kernel that keeps the basis fragments of a row in shared memory --
    rows inside and beyond its 128-entry stash, partial last channel block) against the fp32 exactness path on the same
    inputs: 3e-2 relative."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    pc, neigh, x = _synthetic_layer_problem(1500, 0.2, frames, cin, cout, seed=5, batches=2)
    torch.manual_seed(7)
    layer = PNEConvLayerRotEquiv(9, cin, cout, 32, "mlp_gelu").to(DEV)
    layer.norm_neigh_dist_.fill_(5.0)
    layer.norm_num_neighs_.fill_(1500 / neigh.neighbors_.shape[0])
    dy = torch.randn(1500 * frames, cout, generator=torch.Generator().manual_seed(9)).to(DEV)
    res = []
    for precision in (0, 1):
        layer.precision = precision
        layer.zero_grad()
        xx = x.clone().requires_grad_(True)
        y = layer(pc, pc, xx, neigh)
        y.backward(dy)
        res.append([t.detach().cpu().numpy() for t in (y, xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad,
                                                       layer.proj_biases_.grad)])
    for a, b, name in zip(res[1], res[0], ("y", "dx", "dW", "dA", "dB")):
        err = rel_err(a, b)
        print(cin, cout, frames, name, "bf16 vs fp32 rel err %.2e" % err)
        assert err < 3e-2, name


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("cin,cout,frames", [(32, 32, 2), (1, 32, 2), (160, 72, 4), (64, 128, 1), (40, 8, 3)])
def test_conv_calls_stay_inside_their_buffers(precision, cin, cout, frames):
    """se3_conv_fwd / se3_conv_bwd called directly through the C ABI with every output, saved and workspace buffer
    embedded between 4 KB guard zones: the guards are untouched afterwards (the library only writes what its *_bytes
    functions announce), and the outputs equal those of the layer's own call.  Covers the row-item, one-channel,
    wide-layer (basis stash) and multi-block kernels of both precisions."""
    from se3conv3d_b200 import _lib
    from se3conv3d_b200.custom_ops.functions import make_conv_desc
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    pc, neigh, x = _synthetic_layer_problem(1200, 0.2, frames, cin, cout, seed=21, batches=2)
    torch.manual_seed(22)
    layer = PNEConvLayerRotEquiv(9, cin, cout, 32, "mlp_gelu").to(DEV)
    layer.precision = precision
    layer.norm_neigh_dist_.fill_(5.0)
    layer.norm_num_neighs_.fill_(1200 / neigh.neighbors_.shape[0])
    dy = torch.randn(1200 * frames, cout, generator=torch.Generator().manual_seed(23)).to(DEV)
    xx = x.clone().requires_grad_(True)
    y_ref = layer(pc, pc, xx, neigh)
    y_ref.backward(dy)
    want = [y_ref.detach(), xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad, layer.proj_biases_.grad]

    G = 4096

    def guarded(nbytes):
        nbytes = max(int(nbytes), 16)
        buf = torch.full((nbytes + 2 * G,), 0xA5, dtype=torch.uint8, device=DEV)
        return buf, buf[G:G + nbytes]

    def intact(buf):
        return bool((buf[:G] == 0xA5).all()) and bool((buf[-G:] == 0xA5).all())
    geom = neigh.conv_geometry(pc, pc)
    pa, pb, cw = layer.proj_axes_.detach(), layer.proj_biases_.detach(), layer.conv_weights_.detach()
    d, dref, saved_b, fwd_b, bwd_b, _ = make_conv_desc(geom, cin, cout, 32, 2, precision, float(layer.norm_neigh_dist_),
                                                        float(layer.norm_num_neighs_) / frames, pa, pb, cw)
    L = _lib.lib()
    R = 1200 * frames
    bufs = {name: guarded(nb) for name, nb in (("y", R * cout * 4), ("saved", saved_b), ("fws", fwd_b), ("bws", bwd_b),
                                               ("dx", R * cin * 4), ("dW", cin * 32 * cout * 4), ("dA", 9 * 32 * 4),
                                               ("dB", 32 * 4))}
    p = lambda n: bufs[n][1].data_ptr()
    _lib.check(L.se3_conv_fwd(dref, x.data_ptr(), p("y"), p("saved"), p("fws"), bufs["fws"][1].numel(), _lib.stream()),
               "se3_conv_fwd")
    _lib.check(L.se3_conv_bwd(dref, x.data_ptr(), dy.data_ptr(), p("saved"), p("dx"), p("dW"), p("dA"), p("dB"), p("bws"),
                              bufs["bws"][1].numel(), _lib.stream()), "se3_conv_bwd")
    torch.cuda.synchronize()
    for name, (buf, _) in bufs.items():
        assert intact(buf), "guard zone of %s was written" % name
    got = [bufs["y"][1].view(torch.float32).reshape(R, cout), bufs["dx"][1].view(torch.float32).reshape(R, cin),
           bufs["dW"][1].view(torch.float32).reshape(cin, 32, cout), bufs["dA"][1].view(torch.float32).reshape(9, 32),
           bufs["dB"][1].view(torch.float32).reshape(32)]
    for a, b, name in zip(got, want, ("y", "dx", "dW", "dA", "dB")):
        assert torch.equal(a, b), name


def test_merged_backward_pass_opt_in_matches_fp32_path():
    """The opt-in merged backward pass (SE3_BWD_MERGED=1: basis gradient + per-entry data-gradient contributions in one
    gather pass, k_dx_segsum over the transposed CSR) in a fresh process -- the switch is read once per process --
    against the fp32 exactness path: every gradient within the bf16 tolerance, for 32 / 1 / 24 input channels and
    F_in in {1, 2, 3}."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import test_gpu_parity as tp
from se3conv3d_b200.layers import PNEConvLayerRotEquiv
for cin, cout, frames in ((32, 32, 2), (1, 32, 2), (24, 64, 1), (16, 16, 3)):
    pc, neigh, x = tp._synthetic_layer_problem(1500, 0.2, frames, cin, cout, seed=5, batches=2)
    torch.manual_seed(7)
    layer = PNEConvLayerRotEquiv(9, cin, cout, 32, "mlp_gelu").to("cuda:0")
    layer.norm_neigh_dist_.fill_(5.0)
    layer.norm_num_neighs_.fill_(1500 / neigh.neighbors_.shape[0])
    dy = torch.randn(1500 * frames, cout, generator=torch.Generator().manual_seed(9)).to("cuda:0")
    res = []
    for precision in (0, 1):
        layer.precision = precision
        layer.zero_grad()
        xx = x.clone().requires_grad_(True)
        layer(pc, pc, xx, neigh).backward(dy)
        res.append([t.detach().cpu().numpy() for t in (xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad, layer.proj_biases_.grad)])
    for a, b, name in zip(res[1], res[0], ("dx", "dW", "dA", "dB")):
        err = tp.rel_err(a, b)
        print(cin, cout, frames, name, "%%.2e" %% err)
        assert err < 1e-2, (cin, cout, frames, name, err)
print("MERGED-OK")
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SE3_BWD_MERGED="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MERGED-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("precision,tol", [(0, 1e-4), (1, 3e-2)])
def test_standard_pne_conv_layer_matches_oracle(precision, tol):
    """the non-equivariant PNEConvLayer (layers/PNEConvLayer.py:161-229) on the fused kernels
    against the CPU oracle in float64, outputs and all gradients."""
    from se3conv3d_b200.layers import PNEConvLayer
    from se3conv3d_b200.pc import Pointcloud, BQNeighborhood
    n, r = 3000, 0.15
    pts = torch.rand(n, 3, generator=torch.Generator().manual_seed(11))
    batch = torch.sort(torch.randint(0, 3, (n,), generator=torch.Generator().manual_seed(12)))[0].to(torch.int32)
    pc = Pointcloud(pts.to(DEV), batch.to(DEV))
    neigh = BQNeighborhood(pc, pc, r)
    torch.manual_seed(13)
    layer = PNEConvLayer(3, 24, 40, 32, "mlp_gelu").to(DEV)
    layer.precision = precision
    with torch.no_grad():
        layer.proj_biases_.uniform_(-0.2, 0.2)
    layer.norm_neigh_dist_.fill_(1.0 / r)
    layer.norm_num_neighs_.fill_(n / neigh.neighbors_.shape[0])
    x = torch.randn(n, 24, generator=torch.Generator().manual_seed(14)).to(DEV).requires_grad_(True)
    dy = torch.randn(n, 40, generator=torch.Generator().manual_seed(15)).to(DEV)
    y = layer(pc, pc, x, neigh)
    y.backward(dy)
    c = lambda t: t.detach().cpu().double()
    xo = c(x).requires_grad_(True)
    A, B, W = (c(p).requires_grad_(True) for p in (layer.proj_axes_, layer.proj_biases_, layer.conv_weights_))
    yo = lo.standard_conv_forward(xo, A, B, W, c(pc.pts_), c(pc.pts_), neigh.neighbors_.cpu(),
                                  float(layer.norm_neigh_dist_), float(layer.norm_num_neighs_))
    yo.backward(c(dy))
    assert layer.proj_axes_.grad.shape == (3, 32)
    for got, want, name in ((y, yo, "y"), (x.grad, xo.grad, "dx"), (layer.conv_weights_.grad, W.grad, "dW"),
                            (layer.proj_axes_.grad, A.grad, "dA"), (layer.proj_biases_.grad, B.grad, "dB")):
        err = rel_err(got.detach().cpu().numpy(), want.detach().numpy())
        print("standard layer precision", precision, name, "rel err %.2e" % err)
        assert err < tol, name


def test_config3_classification_shapes_mc_frames():
    """BASELINE config 3 shapes at reduced batch: 1,024-point clouds on a sphere, F=4 Monte-Carlo frames (pca False),
    4 -> 2 frame down-sampling conv: bf16 path vs fp32 path, determinism of the fp32 path."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    from se3conv3d_b200.pc import PointcloudRotEquiv, BQNeighborhood, GridSubSample
    g = torch.Generator().manual_seed(21)
    b_items, n_pts = 8, 1024
    p = torch.randn(b_items * n_pts, 3, generator=g)
    p = p / p.norm(dim=1, keepdim=True) + 0.01 * torch.randn(b_items * n_pts, 3, generator=g)
    batch = torch.arange(b_items).repeat_interleave(n_pts).to(torch.int32)
    cfg4 = {"pca": False, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": 4}
    torch.manual_seed(22)
    pc = PointcloudRotEquiv(p.to(DEV), batch.to(DEV), cfg4)
    assert pc.local_frames_.shape == (b_items * n_pts, 4, 9)
    samp = GridSubSample(pc, 0.1)
    pc2 = PointcloudRotEquiv(samp.__subsample_tensor__(pc.pts_, "avg"), samp.__subsample_tensor__(pc.batch_ids_, "max"),
                             dict(cfg4, n_frames=2))
    for (pin, pout, fi, fo, cin, cout) in ((pc, pc, 4, 4, 16, 32), (pc, pc2, 4, 2, 32, 64)):
        neigh = BQNeighborhood(pin, pout, 0.2)
        torch.manual_seed(23)
        layer = PNEConvLayerRotEquiv(9, cin, cout, 32, "mlp_gelu").to(DEV)
        layer.norm_neigh_dist_.fill_(5.0)
        layer.norm_num_neighs_.fill_(pout.pts_.shape[0] / neigh.neighbors_.shape[0])
        x = torch.randn(pin.pts_.shape[0] * fi, cin, generator=torch.Generator().manual_seed(24)).to(DEV)
        dy = torch.randn(pout.pts_.shape[0] * fo, cout, generator=torch.Generator().manual_seed(25)).to(DEV)
        res = []
        for precision in (0, 0, 1):
            layer.precision = precision
            layer.zero_grad()
            xx = x.clone().requires_grad_(True)
            y = layer(pin, pout, xx, neigh)
            assert y.shape == (pout.pts_.shape[0] * fo, cout)
            y.backward(dy)
            res.append([t.detach().clone() for t in (y, xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad,
                                                     layer.proj_biases_.grad)])
        for a, b in zip(res[0], res[1]):
            assert torch.equal(a, b), "the fp32 path must be deterministic"
        for a, b, name in zip(res[2], res[0], ("y", "dx", "dW", "dA", "dB")):
            err = rel_err(a.cpu().numpy(), b.cpu().numpy())
            print("config3 F=%d->%d" % (fi, fo), name, "bf16 vs fp32 rel err %.2e" % err)
            assert err < 3e-2, name


def test_fused_hierarchy_sampled_frames_config3():
    """Sampled (Monte-Carlo) frames through the fused builder (BASELINE config 3 shapes at reduced batch): the frames
    are the quaternion -> matrix images of the builder's Gaussian draws (re-created from the seed), the level clouds
    equal the per-object grid pooling, and an F = 4 convolution runs on the fused neighbourhood."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    from se3conv3d_b200.pc import build_point_hierarchy, quaternion_to_matrix, Pointcloud, GridSubSample
    g = torch.Generator().manual_seed(51)
    b_items, n_pts, F = 8, 1024, 4
    p = torch.randn(b_items * n_pts, 3, generator=g)
    p = (p / p.norm(dim=1, keepdim=True) + 0.01 * torch.randn(b_items * n_pts, 3, generator=g)).to(DEV)
    batch = torch.arange(b_items).repeat_interleave(n_pts).to(torch.int32).to(DEV)
    cfg = {"pca": False, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": F}
    grids = [0.1, 0.2, 0.4]
    wanted = [(0, 0, 0.1), (0, 1, 0.2), (1, 1, 0.2), (2, 2, 0.4)]
    n = p.shape[0]
    torch.manual_seed(52)
    h, _ = build_point_hierarchy(p, batch, cfg, 0.05, grids, neighborhoods=wanted, n_batches=b_items)
    torch.manual_seed(52)
    q = torch.randn((len(grids) + 2) * n * F * 4, device=DEV)      # the builder's first draw
    off = 0
    pc_ref = Pointcloud(p, batch)
    samp = GridSubSample(pc_ref, 0.05)
    ref_pts = samp.__subsample_tensor__(pc_ref.pts_, "avg")
    for lvl, pc in enumerate(h.pcs_):
        m = pc.pts_.shape[0]
        if lvl == 0:
            assert torch.equal(pc.pts_, ref_pts)
        want = quaternion_to_matrix(q[off * F * 4:(off + m) * F * 4].reshape(-1, 4)).reshape(m, F, 9)
        assert torch.equal(pc.local_frames_, want), lvl
        R = pc.local_frames_.reshape(-1, 3, 3)
        assert float((R @ R.transpose(1, 2) - torch.eye(3, device=DEV)).abs().max()) < 1e-5
        assert float((torch.linalg.det(R) - 1).abs().max()) < 1e-5
        off += m
    sizes = [pc.pts_.shape[0] for pc in h.pcs_]
    assert all(a > b for a, b in zip(sizes, sizes[1:])), sizes
    nb, pc0, pc1 = h.fused_neighborhoods_[1], h.pcs_[0], h.pcs_[1]
    layer = PNEConvLayerRotEquiv(9, 32, 64, 32, "mlp_gelu").to(DEV)
    layer.norm_neigh_dist_.fill_(5.0)
    layer.norm_num_neighs_.fill_(pc1.pts_.shape[0] / nb.neighbors_.shape[0])
    x = torch.randn(pc0.pts_.shape[0] * F, 32, generator=torch.Generator().manual_seed(53)).to(DEV)
    dy = torch.randn(pc1.pts_.shape[0] * F, 64, generator=torch.Generator().manual_seed(54)).to(DEV)
    res = []
    for precision in (0, 1):
        layer.precision = precision
        layer.zero_grad()
        xx = x.clone().requires_grad_(True)
        y = layer(pc0, pc1, xx, nb)
        y.backward(dy)
        res.append([t.detach().cpu().numpy() for t in (y, xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad)])
    for a, b, name in zip(res[1], res[0], ("y", "dx", "dW", "dA")):
        assert rel_err(a, b) < 3e-2, name


def test_config4_scannet_shapes_fixed_axis():
    """BASELINE config 4 shapes: one 150k-point room-like scene, F=1 PCA frames about a fixed up axis, 5-level
    hierarchy through the fused builder; CSR properties at full size and a conv fwd+bwd (bf16 vs fp32)."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    from se3conv3d_b200.pc import build_point_hierarchy
    g = torch.Generator().manual_seed(31)
    n = 150000
    # points on the floor, two walls and a few boxes of an 8 x 6 x 3 m room
    u = torch.rand(n, 3, generator=g)
    which = torch.randint(0, 4, (n,), generator=g)
    p = torch.stack((u[:, 0] * 8, u[:, 1] * 6, u[:, 2] * 3), 1)
    p[which == 0, 2] = 0.0
    p[which == 1, 0] = 0.0
    p[which == 2, 1] = 0.0
    p[which == 3] = p[which == 3] * torch.tensor([0.15, 0.2, 0.3]) + torch.tensor([3.0, 2.0, 0.0])
    p = (p + 0.004 * torch.randn(n, 3, generator=g)).to(torch.float32)
    batch = torch.zeros(n, dtype=torch.int32)
    cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": 2, "n_frames": 1}
    grids = [0.2, 0.4, 0.8, 1.6]
    wanted = [(0, 0, 0.2), (0, 1, 0.2), (1, 1, 0.4), (2, 1, 0.8), (4, 3, 3.2)]
    h, _ = build_point_hierarchy(p.to(DEV), batch.to(DEV), cfg, 0.1, grids, neighborhoods=wanted, n_batches=1)
    sizes = [pc.pts_.shape[0] for pc in h.pcs_]
    assert len(sizes) == 5 and all(a > b for a, b in zip(sizes, sizes[1:])), sizes
    for pc in h.pcs_:
        assert pc.local_frames_.shape == (pc.pts_.shape[0], 1, 9)
        R = pc.local_frames_.reshape(-1, 3, 3)
        assert float((R @ R.transpose(1, 2) - torch.eye(3, device=DEV)).abs().max()) < 1e-4   # orthonormal frames
    for nb, (s, t, r) in zip(h.fused_neighborhoods_, wanted):
        ends = nb.start_ids_.to(torch.int64)
        assert bool((ends[1:] >= ends[:-1]).all())
        pairs = nb.neighbors_
        assert int(ends[-1]) == pairs.shape[0]
        d = (h.pcs_[s].pts_[pairs[:, 1]] - h.pcs_[t].pts_[pairs[:, 0]]).norm(dim=1)
        assert float(d.max()) < r * (1 + 1e-5)
        assert bool((pairs[1:, 0] >= pairs[:-1, 0]).all())          # grouped by sample
    pc, nb = h.pcs_[0], h.fused_neighborhoods_[0]
    torch.manual_seed(32)
    layer = PNEConvLayerRotEquiv(9, 32, 32, 32, "mlp_gelu").to(DEV)
    layer.norm_neigh_dist_.fill_(5.0)
    layer.norm_num_neighs_.fill_(pc.pts_.shape[0] / nb.neighbors_.shape[0])
    x = torch.randn(pc.pts_.shape[0], 32, generator=torch.Generator().manual_seed(33)).to(DEV)
    dy = torch.randn(pc.pts_.shape[0], 32, generator=torch.Generator().manual_seed(34)).to(DEV)
    res = []
    for precision in (0, 1):
        layer.precision = precision
        layer.zero_grad()
        xx = x.clone().requires_grad_(True)
        y = layer(pc, pc, xx, nb)
        y.backward(dy)
        res.append([t.detach().cpu().numpy() for t in (y, xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad,
                                                       layer.proj_biases_.grad)])
    for a, b, name in zip(res[1], res[0], ("y", "dx", "dW", "dA", "dB")):
        err = rel_err(a, b)
        print("config4 F=1", name, "bf16 vs fp32 rel err %.2e" % err)
        assert err < 3e-2, name


def test_resnetformer_block_forward_backward():
    """Row f2: a ResNetFormer block around the fused conv trains: finite outputs / gradients for every parameter,
    and with gamma = 1e-6 the block output is the skip path within 1e-4."""
    from se3conv3d_b200.layers import ResNetFormer, BatchNormPC, PNEConvLayerRotEquivFactory
    pc, neigh, x = _synthetic_layer_problem(2000, 0.18, 2, 32, 32, seed=41, batches=3)
    torch.manual_seed(42)
    fact = PNEConvLayerRotEquivFactory(9, 32, "mlp_gelu")
    blk = ResNetFormer(32, 48, fact, BatchNormPC, 0.0).to(DEV)
    blk.spatial_conv_.precision = 1
    blk.spatial_conv_.norm_neigh_dist_.fill_(1 / 0.18)
    blk.spatial_conv_.norm_num_neighs_.fill_(2000 / neigh.neighbors_.shape[0])
    xx = x.clone().requires_grad_(True)
    y = blk(pc, xx, neigh)
    assert y.shape == (4000, 48) and bool(torch.isfinite(y).all())
    assert rel_err(y.detach().cpu().numpy(), blk.skip_conv_(xx).detach().cpu().numpy()) < 1e-4
    y.square().mean().backward()
    for name, p in blk.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name
    assert float(blk.spatial_conv_.conv_weights_.grad.abs().max()) > 0.0


@pytest.mark.parametrize("precision,tol", [(0, 1e-4), (1, 3e-2)])
def test_layer_ragged_rows_and_empty_neighbourhoods(precision, tol):
    """Ragged CSR: output points without any neighbour (empty rows -> exact zeros), rows longer than one 32-entry
    chunk and longer than the 32-edge id prefetch, two batch items of very different size; plus a neighbourhood
    with no edge at all (zero output, zero gradients)."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    from se3conv3d_b200.pc import PointcloudRotEquiv, BQNeighborhood
    g = torch.Generator().manual_seed(51)
    cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": 2}
    dense = torch.rand(900, 3, generator=g) * 0.3                       # item 0: dense blob (rows of ~100+ edges)
    sparse = torch.rand(60, 3, generator=g) * 4.0 + 10.0                # item 1: isolated points
    src_pts = torch.cat((dense, sparse))
    src_b = torch.cat((torch.zeros(900), torch.ones(60))).to(torch.int32)
    dst_pts = torch.cat((torch.rand(200, 3, generator=g) * 0.3, torch.rand(40, 3, generator=g) * 4.0 + 30.0))
    dst_b = torch.cat((torch.zeros(200), torch.ones(40))).to(torch.int32)
    torch.manual_seed(52)
    pc_in = PointcloudRotEquiv(src_pts.to(DEV), src_b.to(DEV), cfg)
    pc_out = PointcloudRotEquiv(dst_pts.to(DEV), dst_b.to(DEV), cfg)
    neigh = BQNeighborhood(pc_in, pc_out, 0.12)
    ends = neigh.start_ids_.cpu().numpy()
    counts = np.diff(ends, prepend=0)
    assert counts[200:].max() == 0 and counts[:200].max() > 40          # empty rows and rows > 32 edges
    torch.manual_seed(53)
    layer = PNEConvLayerRotEquiv(9, 32, 32, 32, "mlp_gelu").to(DEV)
    layer.precision = precision
    layer.norm_neigh_dist_.fill_(1 / 0.12)
    layer.norm_num_neighs_.fill_(240 / max(neigh.neighbors_.shape[0], 1))
    x = torch.randn(960 * 2, 32, generator=torch.Generator().manual_seed(54)).to(DEV).requires_grad_(True)
    dy = torch.randn(240 * 2, 32, generator=torch.Generator().manual_seed(55)).to(DEV)
    y = layer(pc_in, pc_out, x, neigh)
    y.backward(dy)
    assert float(y[400:].abs().max()) == 0.0                             # rows without neighbours
    assert float(x.grad[1800:].abs().max()) == 0.0                       # sources nobody gathers
    c = lambda t: t.detach().cpu().double()
    xo = c(x).requires_grad_(True)
    A, B, W = (c(p).requires_grad_(True) for p in (layer.proj_axes_, layer.proj_biases_, layer.conv_weights_))
    yo = lo.conv_forward(xo, A, B, W, c(pc_in.pts_), c(pc_out.pts_), c(pc_in.local_frames_), c(pc_out.local_frames_),
                         neigh.neighbors_.cpu(), float(layer.norm_neigh_dist_), float(layer.norm_num_neighs_))
    yo.backward(c(dy))
    for got, want, name in ((y, yo, "y"), (x.grad, xo.grad, "dx"), (layer.conv_weights_.grad, W.grad, "dW"),
                            (layer.proj_axes_.grad, A.grad, "dA"), (layer.proj_biases_.grad, B.grad, "dB")):
        assert rel_err(got.detach().cpu().numpy(), want.detach().numpy()) < tol, name
    # no edge at all
    far = PointcloudRotEquiv((dst_pts + 100.0).to(DEV), dst_b.to(DEV), cfg)
    none = BQNeighborhood(pc_in, far, 0.05)
    assert none.neighbors_.shape[0] == 0
    layer.zero_grad()
    x2 = x.detach().clone().requires_grad_(True)
    y0 = layer(pc_in, far, x2, none)
    y0.backward(dy)
    assert float(y0.abs().max()) == 0.0 and float(x2.grad.abs().max()) == 0.0
    assert float(layer.conv_weights_.grad.abs().max()) == 0.0 and float(layer.proj_axes_.grad.abs().max()) == 0.0


def test_segment_pool_is_differentiable():
    """Grid pooling of a feature tensor that requires grad (the reference uses torch_scatter's scatter_mean /
    scatter_max, pc/GridSubSample.py:69-72): values and gradients against index_add / scatter_reduce."""
    from se3conv3d_b200.pc import Pointcloud, GridSubSample
    pts, b = cloud(4000, 3, 21)
    pc = Pointcloud(pts.to(DEV), b.to(DEV))
    samp = GridSubSample(pc, 0.12)
    ids, m = samp.grid_.cell_ids_, samp.grid_.num_used_cells_
    feats = torch.randn(4000, 7, generator=torch.Generator().manual_seed(1)).to(DEV)
    w = torch.randn(m, 7, generator=torch.Generator().manual_seed(2)).to(DEV)
    for method, red in (("avg", "mean"), ("max", "amax")):
        x = feats.clone().requires_grad_(True)
        out = samp.__subsample_tensor__(x, method)
        assert out.requires_grad
        (out * w).sum().backward()
        x2 = feats.clone().requires_grad_(True)
        ref = torch.zeros(m, 7, device=DEV).scatter_reduce(0, ids[:, None].expand(-1, 7), x2, reduce=red, include_self=False)
        (ref * w).sum().backward()
        assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
        assert rel_err(x.grad.cpu().numpy(), x2.grad.cpu().numpy()) < 1e-6, method


@pytest.mark.parametrize("case", ["matrix", "quaternion", "softmax", "basis16"])
def test_layer_variants_match_reference_golden(case):
    """The configurations outside the fused kernels -- 'matrix' / 'quaternion' relative rotations
    (pc/RotationFunctions.py:593-600), the mlp_softmax basis (layers/PNEConvLayer.py:97), 16 basis functions -- through
    the composed GPU path (reference statement sequence over this package's FeatBasisProj op) against the reference's
    own float64 forward / backward (tests/golden/gen_variant_golden.py): <= 1e-4 on y and all four gradients."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv, PNEConvLayerRotEquivFactory
    from se3conv3d_b200.pc import BQNeighborhood
    g = dict(np.load(os.path.join(GOLDEN, "variant_%s.npz" % case)))
    t = lambda k: torch.from_numpy(g[k]).to(DEV)
    f = g["frames"].shape[1]
    pc = types.SimpleNamespace(pts_=t("pts"), local_frames_=t("frames"), n_frames_=f)
    neigh = BQNeighborhood.__new__(BQNeighborhood)
    neigh.neighbors_, neigh.start_ids_, neigh.conv_geometry_cache_ = t("neighbors"), t("ends"), {}
    cin, k, cout = g["conv_weights"].shape
    fac = PNEConvLayerRotEquivFactory(int(g["dims"]), k, str(g["pne"]), p_rel_rot=str(g["rel"]))   # sets rel_rot_type
    try:
        layer = fac.create_conv_layer(cin, cout).to(DEV)
        with torch.no_grad():
            layer.proj_axes_.copy_(t("proj_axes"))
            layer.proj_biases_.copy_(t("proj_biases"))
            layer.conv_weights_.copy_(t("conv_weights"))
            layer.norm_neigh_dist_.fill_(float(g["norm_neigh_dist"]))
            layer.norm_num_neighs_.fill_(float(g["norm_num_neighs"]))
        x = t("x").clone().requires_grad_(True)
        y = layer(pc, pc, x, neigh)
        (y * t("dy")).sum().backward()
        got = (y, x.grad, layer.conv_weights_.grad, layer.proj_axes_.grad, layer.proj_biases_.grad)
        for name, a, key in zip(("y", "dx", "dW", "dA", "dB"), got, ("y", "dx", "dW", "dA", "dB")):
            m = lo.err_metrics(a.detach().cpu().numpy(), g[key])
            print(case, name, "max/max %.2e relL2 %.2e p99.9 %.2e" % m)
            assert max(m) < 1e-4, (case, name, m)
    finally:
        PNEConvLayerRotEquiv.rel_rot_type = "6D"


@pytest.mark.parametrize("k", [16, 33, 48, 64])
def test_knn_up_to_64_and_cross_cloud(k):
    """k up to 64 (two ranks per lane; the reference op's limit, knn_query.cu:135-197) and the cross-cloud query
    (pc/KnnNeighborhood.py:78-84) against exact brute-force distances; ties by distance excluded through the sorted
    distance rows."""
    from se3conv3d_b200.pc import Pointcloud, KnnNeighborhood
    pts, b = cloud(3000, 3, 41, scale=(1.0, 0.7, 0.4))
    pc = Pointcloud(pts.to(DEV), b.to(DEV))
    nbh = KnnNeighborhood(pc, pc, k, p_keep_empty=True)
    tab = nbh.knn_table_.cpu().numpy()
    P, B = pts.numpy().astype(np.float64), b.numpy()
    full = ((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)
    full[B[:, None] != B[None, :]] = np.inf
    ref = np.sort(full, axis=1)[:, :k]
    got = np.where(tab >= 0, full[np.arange(3000)[:, None], np.maximum(tab, 0)], np.inf)
    np.testing.assert_allclose(np.sort(got, 1), ref, rtol=1e-5, atol=1e-12)
    assert nbh.neighbors_.shape == (3000 * k, 2) and int(nbh.start_ids_[-1]) == 3000 * k
    # cross-cloud: 500 samples against the 3000 sources of their batch item
    spts, sb = cloud(500, 3, 43, scale=(1.0, 0.7, 0.4))
    spc = Pointcloud(spts.to(DEV), sb.to(DEV))
    x = KnnNeighborhood(pc, spc, k)
    xt = x.knn_table_.cpu().numpy()
    S = spts.numpy().astype(np.float64)
    cross = ((S[:, None, :] - P[None, :, :]) ** 2).sum(-1)
    cross[sb.numpy()[:, None] != B[None, :]] = np.inf
    refx = np.sort(cross, axis=1)[:, :k]
    gotx = np.where(xt >= 0, cross[np.arange(500)[:, None], np.maximum(xt, 0)], np.inf)
    np.testing.assert_allclose(gotx, refx, rtol=1e-5, atol=1e-12)          # already in ascending order
    assert x.neighbors_.shape[0] == int(x.start_ids_[-1]) and x.neighbors_.shape[1] == 2


def test_weight_layout_cache_follows_the_parameter():
    """The per-layer cache of the bf16 weight layouts (se3_conv_desc.weight_cache): a second call with unchanged weights
    skips the conversion and is bit-identical; after an in-place update (optimiser step) or a replaced parameter tensor
    the layouts are rebuilt -- results equal those of a fresh layer with the same weights."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    pc, neigh, x = _synthetic_layer_problem(2048, 0.16, 2, 32, 64)
    torch.manual_seed(2)
    layer = PNEConvLayerRotEquiv(9, 32, 64, 32, "mlp_gelu").to(DEV)
    layer.precision = 1
    layer.norm_neigh_dist_.fill_(1 / 0.16)
    layer.norm_num_neighs_.fill_(2048 / neigh.neighbors_.shape[0])

    def run(l):
        xx = x.clone().requires_grad_(True)
        y = l(pc, pc, xx, neigh)
        y.square().mean().backward()
        return y.detach().clone(), xx.grad.clone(), l.conv_weights_.grad.clone()
    a = run(layer)
    assert layer._wcache.key is not None
    layer.zero_grad()
    b = run(layer)                                   # cached layouts
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    with torch.no_grad():
        layer.conv_weights_.mul_(1.5)                # in-place update: version bump
    layer.zero_grad()
    c = run(layer)
    fresh = PNEConvLayerRotEquiv(9, 32, 64, 32, "mlp_gelu").to(DEV)
    fresh.precision = 1
    fresh.load_state_dict(layer.state_dict())
    d = run(fresh)
    assert all(torch.equal(u, v) for u, v in zip(c, d))
    assert not torch.equal(a[0], c[0])
