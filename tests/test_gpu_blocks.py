"""GPU parity of the glue around the convolution against the REFERENCE run on the CPU
(tests/golden/gen_block_golden.py): the ResNetFormer block (BatchNormPC -> conv -> gamma-skip -> BatchNormPC -> Linear x2
-> GELU -> Linear -> gamma-skip; layers/ResNetFormer.py:52-91) forward, input gradient and every parameter gradient, and
the frame / global pooling helpers (pc/PointcloudRotEquiv.py:195-286).  The gamma-skip (+ drop path) and the poolings
run as single kernels of csrc/block_ops.cu."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from fpn_fixture import reinit_by_name
from oracle import layer_oracle as lo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 8}, "fixed_axis": False, "n_frames": 2}


def _cloud(g):
    from se3conv3d_b200.pc import PointcloudRotEquiv, BQNeighborhood
    pc = PointcloudRotEquiv.__new__(PointcloudRotEquiv)
    pc.pts_with_grads_, pc.batch_size_host_, pc._batch_size = False, 3, None
    pc.pts_ = torch.from_numpy(g["pts"]).to(DEV)
    pc.batch_ids_ = torch.from_numpy(g["batch"]).to(DEV)
    pc.neigh_cache_, pc.local_frames_pca_cache_ = {}, {}
    pc.local_frames_config_, pc.standard_knn_, pc.ref_frames_pts = CFG, False, None
    pc.local_frames_ = torch.from_numpy(g["frames"]).to(DEV).contiguous()
    pc.n_frames_, pc._batch_ids_frames = 2, None
    nb = BQNeighborhood(pc, pc, 0.4)
    assert np.array_equal(nb.start_ids_.cpu().numpy(), g["ends"])
    canon = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
    assert np.array_equal(canon(nb.neighbors_.cpu().numpy()), canon(g["neighbors"].astype(np.int64)))
    return pc, nb


@pytest.mark.parametrize("name,cin,cout", [("same", 16, 16), ("widen", 16, 24)])
def test_resnetformer_block_matches_reference(name, cin, cout):
    from se3conv3d_b200.layers import ResNetFormer, BatchNormPC, PNEConvLayerRotEquivFactory
    g = dict(np.load(os.path.join(GOLDEN, "block_resnetformer.npz")))
    pc, nb = _cloud(g)
    blk = ResNetFormer(cin, cout, PNEConvLayerRotEquivFactory(9, 32, "mlp_gelu"), BatchNormPC, 0.0)
    assert sorted(k for k, _ in blk.named_parameters()) == [str(s) for s in g[name + "_param_names"]]
    reinit_by_name(blk)
    blk = blk.to(DEV)
    blk.spatial_conv_.norm_neigh_dist_.fill_(1.0 / 0.4)
    blk.spatial_conv_.norm_num_neighs_.fill_(g["pts"].shape[0] / g["neighbors"].shape[0])
    blk.train()
    x = torch.from_numpy(g[name + "_x"]).to(DEV).requires_grad_(True)
    y = blk(pc, x, nb)
    (y * torch.from_numpy(g[name + "_dy"]).to(DEV)).sum().backward()
    worst = 0.0
    for tag, got, ref in [("y", y, g[name + "_y"]), ("dx", x.grad, g[name + "_dx"])] + \
                         [("d " + k, p.grad, g[name + "_grad_" + k]) for k, p in blk.named_parameters()]:
        m = lo.err_metrics(got.detach().cpu().numpy(), ref)
        worst = max(worst, m[0], m[1])
        assert m[0] < 1e-4 and m[1] < 1e-4, (name, tag, m)
    print(name, "block: worst max/max | relL2 over output and %d gradients: %.2e" % (len(list(blk.parameters())) + 1, worst))


def test_frame_and_global_pooling_match_reference():
    g = dict(np.load(os.path.join(GOLDEN, "block_resnetformer.npz")))
    pc, _ = _cloud(g)
    x = torch.from_numpy(g["pool_x"]).float().to(DEV)
    for m in ("avg", "sum", "max", "min"):
        assert lo.err_metrics(pc.feature_pooling(x, m).cpu().numpy(), g["frame_" + m])[0] < 1e-6, m
        assert lo.err_metrics(pc.global_pooling(x, m).cpu().numpy(), g["global_" + m])[0] < 1e-6, m
    got = pc.global_pooling_specific_feature_pooling(x, "avg", "max")
    assert lo.err_metrics(got.cpu().numpy(), g["global_specific_avg_max"])[0] < 1e-6
    up = pc.global_upsample(torch.arange(21, dtype=torch.float32, device=DEV).reshape(3, 7))
    np.testing.assert_array_equal(up.cpu().numpy(), g["upsample"].astype(np.float32))
    # gradients of the pooling kernels against autograd over the plain tensor formulation
    n, f = pc.pts_.shape[0], 2
    w = torch.randn(n, 7, generator=torch.Generator().manual_seed(3)).to(DEV)
    wb = torch.randn(3, 7, generator=torch.Generator().manual_seed(4)).to(DEV)
    ids = pc.batch_ids_considering_frames_.to(torch.int64)
    for m in ("avg", "sum", "max", "min"):
        a = x.clone().requires_grad_(True)
        (pc.feature_pooling(a, m) * w).sum().backward()
        b = x.clone().requires_grad_(True)
        r = b.reshape(n, f, 7)
        ref = {"avg": r.mean(1), "sum": r.sum(1), "max": r.max(1)[0], "min": r.min(1)[0]}[m]
        (ref * w).sum().backward()
        assert lo.err_metrics(a.grad.cpu().numpy(), b.grad.cpu().numpy())[0] < 1e-6, m
    for m in ("avg", "sum"):
        a = x.clone().requires_grad_(True)
        (pc.global_pooling(a, m) * wb).sum().backward()
        b = x.clone().requires_grad_(True)
        s = torch.zeros(3, 7, device=DEV).index_add(0, ids, b)
        if m == "avg":
            s = s / torch.bincount(ids, minlength=3).float()[:, None]
        (s * wb).sum().backward()
        assert lo.err_metrics(a.grad.cpu().numpy(), b.grad.cpu().numpy())[0] < 1e-6, m


def test_gamma_skip_with_drop_path_mask():
    """se3_gamma_skip_fwd / _bwd with an explicit per-item keep mask against the statement sequence of
    SkipConnection + DropPathPC (x * gamma, / keep_prob, * mask[batch of the row], + y) and its autograd gradients."""
    from se3conv3d_b200.custom_ops import GammaSkip
    gen = torch.Generator().manual_seed(9)
    for rows_pts, f, c in ((700, 2, 32), (513, 1, 300), (40, 4, 7)):
        items = torch.sort(torch.randint(0, 5, (rows_pts,), generator=gen))[0].to(torch.int32).to(DEV)
        keep = 0.6
        mask = torch.floor(keep + torch.rand(5, generator=gen)).to(DEV)
        scale = (mask / keep).contiguous()
        x = torch.randn(rows_pts * f, c, generator=gen).to(DEV)
        y = torch.randn(rows_pts * f, c, generator=gen).to(DEV)
        gamma = torch.randn(1, c, generator=gen).to(DEV)
        w = torch.randn(rows_pts * f, c, generator=gen).to(DEV)
        a = [t.clone().requires_grad_(True) for t in (x, y, gamma)]
        (GammaSkip.apply(a[0], a[1], a[2], scale, items, f) * w).sum().backward()
        b = [t.clone().requires_grad_(True) for t in (x, y, gamma)]
        row_items = torch.repeat_interleave(items.to(torch.int64), f)
        ref = (b[0] * b[2]).div(keep) * mask[row_items].reshape(-1, 1) + b[1]
        (ref * w).sum().backward()
        for got, want in zip(a, b):
            assert lo.err_metrics(got.grad.cpu().numpy(), want.grad.cpu().numpy())[1] < 1e-5
        # no drop path
        out = GammaSkip.apply(x, y, gamma, None, None, f)
        assert lo.err_metrics(out.cpu().numpy(), (x * gamma + y).cpu().numpy())[0] < 1e-6
