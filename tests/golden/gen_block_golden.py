"""Generates tests/golden/block_*.npz: the REFERENCE's ResNetFormer block (layers/ResNetFormer.py:52-91 with
BatchNormPC, SkipConnection, DropPathPC and PNEConvLayerRotEquiv inside) forward + autograd backward in float64 on the
CPU, and the reference's frame / global pooling helpers (pc/PointcloudRotEquiv.py:195-286) on the same cloud.

    python tests/golden/gen_block_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle.shims_cpu.point_cloud_lib_ops as cpu_ops  # noqa: E402
sys.modules["point_cloud_lib_ops"] = cpu_ops
from oracle.ref_import import import_reference  # noqa: E402
from fpn_fixture import reinit_by_name  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 8}, "fixed_axis": False, "n_frames": 2}


def main():
    pclib = import_reference()
    refmod = sys.modules["point_cloud_lib.layers.PNEConvLayerRotEquiv"]

    class ScatterFeatBasisProj:
        @staticmethod
        def apply(basis, feats, neighbors, ends):
            return cpu_ops.feat_basis_proj(basis, feats, neighbors, ends)
    refmod.FeatBasisProj = ScatterFeatBasisProj
    gen = torch.Generator().manual_seed(21)
    n = 90
    pts = torch.rand(n, 3, generator=gen)
    b = torch.sort(torch.randint(0, 3, (n,), generator=gen))[0].to(torch.int32)
    torch.manual_seed(5)
    pc = pclib.pc.PointcloudRotEquiv(pts, b, CFG)
    neigh = pclib.pc.BQNeighborhood(pc, pc, 0.4)
    out = {"pts": pts.numpy(), "batch": b.numpy(), "frames": pc.local_frames_.numpy(), "neighbors": neigh.neighbors_.numpy(),
           "ends": neigh.start_ids_.numpy()}
    for name, (cin, cout) in {"same": (16, 16), "widen": (16, 24)}.items():
        fac = pclib.layers.PNEConvLayerRotEquivFactory(9, 32, "mlp_gelu")
        blk = pclib.layers.ResNetFormer(cin, cout, fac, pclib.layers.BatchNormPC, 0.0)
        reinit_by_name(blk)
        blk = blk.double()
        blk.spatial_conv_.norm_neigh_dist_ = torch.tensor(1.0 / 0.4, dtype=torch.float64)
        blk.spatial_conv_.norm_num_neighs_ = torch.tensor(n / neigh.neighbors_.shape[0], dtype=torch.float64)
        blk.train()
        x = torch.randn(n * 2, cin, generator=gen)
        dy = torch.randn(n * 2, cout, generator=gen)
        pcd = pclib.pc.PointcloudRotEquiv.__new__(pclib.pc.PointcloudRotEquiv)
        pcd.__dict__.update(pc.__dict__)
        pcd.pts_, pcd.local_frames_ = pc.pts_.double(), pc.local_frames_.double()
        refmod.PNEConvLayerRotEquiv.empty_rot_tenors_cache()
        xr = x.double().clone().requires_grad_(True)
        y = blk(pcd, xr, neigh)
        (y * dy.double()).sum().backward()
        out.update({name + "_x": x.numpy(), name + "_dy": dy.numpy(), name + "_y": y.detach().numpy(), name + "_dx": xr.grad.numpy()})
        names = sorted(k for k, _ in blk.named_parameters())
        out[name + "_param_names"] = np.array(names)
        for k, p in blk.named_parameters():
            out[name + "_grad_" + k] = p.grad.numpy()
        print(name, "y", tuple(y.shape), "params", len(names))
    # pooling helpers of the reference on a feature matrix
    feats = torch.randn(n * 2, 7, generator=gen).double()
    out["pool_x"] = feats.numpy()
    for m in ("avg", "sum", "max", "min"):
        out["frame_" + m] = pc.feature_pooling(feats, m).numpy()
        out["global_" + m] = pc.global_pooling(feats, m).numpy()
    out["global_specific_avg_max"] = pc.global_pooling_specific_feature_pooling(feats, "avg", "max").numpy()
    out["upsample"] = pc.global_upsample(torch.arange(3 * 7, dtype=torch.float64).reshape(3, 7)).numpy()
    np.savez_compressed(os.path.join(OUT, "block_resnetformer.npz"), **out)
    print("written")


if __name__ == "__main__":
    main()
