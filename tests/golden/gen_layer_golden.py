"""Generates tests/golden/layer_*.npz and frames_*.npz by running the REFERENCE's own Python
(/root/reference, imported through oracle/ref_import.py with the torch_scatter / torch_cluster
shims) on seeded synthetic inputs, on the CPU of the build container.

    python tests/golden/gen_layer_golden.py

What runs from the reference, unmodified:
  PNEConvLayerRotEquiv.get_rot_tenors + __compute_convolution__  (layers/PNEConvLayerRotEquiv.py:61-216)
  sample_reference_frames_pca, sample_reference_frames           (pc/RotationFunctions.py:307-508)
The one substitution: the CUDA-only FeatBasisProj op is replaced by its scatter formulation
  T = scatter_add(h[:,None,:] * x[src][:,:,None], row)
(the reference has no CPU implementation of that op).  Gradients come from
torch.autograd over the reference forward.  float64 copies are stored too (tight oracle checks).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def brute_radius(pts_in, pts_out, b_in, b_out, r):
    d = torch.cdist(pts_out.double(), pts_in.double())
    ok = (d < r) & (b_out[:, None] == b_in[None, :])
    rows, cols = torch.nonzero(ok, as_tuple=True)
    nb = torch.stack((rows, cols), 1)
    counts = torch.bincount(rows, minlength=pts_out.shape[0])
    return nb, torch.cumsum(counts, 0).to(torch.int32)


def main():
    import_reference()
    refmod = sys.modules["point_cloud_lib.layers.PNEConvLayerRotEquiv"]
    from point_cloud_lib.pc import sample_reference_frames_pca, sample_reference_frames
    from point_cloud_lib.pc.RotationFunctions import quaternion_to_matrix

    class ScatterFeatBasisProj:
        @staticmethod
        def apply(basis, feats, neighbors, ends):
            t = torch.zeros((ends.shape[0], feats.shape[1], basis.shape[1]), dtype=feats.dtype)
            return t.index_add(0, neighbors[:, 0].long(), feats[neighbors[:, 1].long()][:, :, None] * basis[:, None, :])

    refmod.FeatBasisProj = ScatterFeatBasisProj

    cases = {
        # name: (N, M or None for pc_in is pc_out, batches, F_in, F_out, Cin, Cout, radius, pne)
        "same_f2": (96, None, 2, 2, 2, 8, 16, 0.45, "mlp_gelu"),
        "same_f1": (64, None, 1, 1, 1, 3, 8, 0.5, "mlp_gelu"),
        "down_f4f2": (120, 40, 2, 4, 2, 16, 8, 0.6, "mlp_gelu"),
        "same_f2_relu": (48, None, 1, 2, 2, 1, 8, 0.6, "mlp_relu"),
    }
    for name, (n, m, nb_batches, fi, fo, cin, cout, radius, pne) in cases.items():
        gen = torch.Generator().manual_seed(sum(map(ord, name)))
        pts_in = torch.rand(n, 3, generator=gen)
        b_in = torch.sort(torch.randint(0, nb_batches, (n,), generator=gen))[0]
        q = torch.randn(n * fi, 4, generator=gen)
        fr_in = quaternion_to_matrix(q / q.norm(dim=1, keepdim=True)).reshape(n, fi, 9)
        if m is None:
            pts_out, b_out, fr_out, m_ = pts_in, b_in, fr_in, n
            assert fi == fo
        else:
            m_ = m
            pts_out = torch.rand(m, 3, generator=gen)
            b_out = torch.sort(torch.randint(0, nb_batches, (m,), generator=gen))[0]
            q = torch.randn(m * fo, 4, generator=gen)
            fr_out = quaternion_to_matrix(q / q.norm(dim=1, keepdim=True)).reshape(m, fo, 9)
        nb, ends = brute_radius(pts_in, pts_out, b_in, b_out, radius)
        # make sure the last output point has neighbours (the reference sizes y by the largest populated row)
        assert int(nb[:, 0].max()) == m_ - 1, name
        x = torch.randn(n * fi, cin, generator=gen)
        dy = torch.randn(m_ * fo, cout, generator=gen)
        torch.manual_seed(1234)
        layer = refmod.PNEConvLayerRotEquiv(9, cin, cout, 32, pne)
        with torch.no_grad():
            layer.proj_biases_.copy_(0.1 * torch.randn(32, generator=gen))
        layer.norm_neigh_dist_ = torch.tensor(1.0 / radius, dtype=torch.float32)
        layer.norm_num_neighs_ = torch.tensor(m_ / nb.shape[0], dtype=torch.float32)
        out = {}
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            refmod.PNEConvLayerRotEquiv.empty_rot_tenors_cache()
            pc_in = types.SimpleNamespace(pts_=pts_in.to(dt), local_frames_=fr_in.to(dt), n_frames_=fi)
            pc_out = pc_in if m is None else types.SimpleNamespace(pts_=pts_out.to(dt), local_frames_=fr_out.to(dt),
                                                                   n_frames_=fo)
            neigh = types.SimpleNamespace(neighbors_=nb.clone(), start_ids_=ends.clone())
            lay = refmod.PNEConvLayerRotEquiv(9, cin, cout, 32, pne).to(dt)
            lay.load_state_dict({k: v.to(dt) for k, v in layer.state_dict().items()})
            lay.norm_neigh_dist_ = layer.norm_neigh_dist_.to(dt)
            lay.norm_num_neighs_ = layer.norm_num_neighs_.to(dt)
            xr = x.to(dt).clone().detach().requires_grad_(True)
            y = lay(pc_in, pc_out, xr, neigh)
            assert y.shape[0] == m_ * fo
            (y * dy.to(dt)).sum().backward()
            rt = refmod.PNEConvLayerRotEquiv.get_rot_tenors(pc_in, pc_out, neigh, lay.norm_neigh_dist_)
            out.update({"y_" + tag: y.detach().numpy(), "dx_" + tag: xr.grad.numpy(),
                        "dW_" + tag: lay.conv_weights_.grad.numpy(), "dA_" + tag: lay.proj_axes_.grad.numpy(),
                        "dB_" + tag: lay.proj_biases_.grad.numpy()})
            if tag == "f64":
                out["g_sorted"] = rt["rel_pts_rel_orient"].numpy()
                out["nb_expanded"] = rt["neighbs"].numpy()
                out["ends_expanded"] = rt["neighbs_start_ids"].numpy()
        np.savez_compressed(
            os.path.join(OUT, "layer_%s.npz" % name), pts_in=pts_in.numpy(), pts_out=pts_out.numpy(),
            frames_in=fr_in.numpy(), frames_out=fr_out.numpy(), neighbors=nb.numpy(), ends=ends.numpy(),
            batch_in=b_in.numpy(), batch_out=b_out.numpy(),
            x=x.numpy(), dy=dy.numpy(), proj_axes=layer.proj_axes_.detach().numpy(),
            proj_biases=layer.proj_biases_.detach().numpy(), conv_weights=layer.conv_weights_.detach().numpy(),
            norm_neigh_dist=np.float32(layer.norm_neigh_dist_), norm_num_neighs=np.float32(layer.norm_num_neighs_),
            pne=np.array(pne), same=np.array(m is None), **out)
        print(name, "E =", nb.shape[0], "y", tuple(out["y_f32"].shape))

    # ---- frames -------------------------------------------------------------------------------
    gen = torch.Generator().manual_seed(77)
    n, k = 200, 16
    pts = torch.rand(n, 3, generator=gen) * torch.tensor([1.0, 0.6, 0.3])
    d = torch.cdist(pts.double(), pts.double())
    knn = torch.topk(d, k, dim=1, largest=False)[1]
    knn[5, 10:] = -1  # missing neighbours -> self loops
    nbr = torch.stack((torch.arange(n)[:, None].expand(n, k).reshape(-1), knn.reshape(-1)), 1)
    res = {"pts": pts.numpy(), "knn": knn.numpy().astype(np.int32)}
    for axis, tag in ((False, "none"), (2, "axis2"), (1, "axis1")):
        neigh = types.SimpleNamespace(neighbors_=nbr.clone(), k_=k)
        res["frames_" + tag] = sample_reference_frames_pca(pts.clone(), neigh, axis_fixed=axis, device="cpu").numpy()
        res["frames64_" + tag] = sample_reference_frames_pca(
            pts.clone().double(), types.SimpleNamespace(neighbors_=nbr.clone(), k_=k), axis_fixed=axis,
            device="cpu").numpy()
    torch.manual_seed(99)
    res["mc_frames"] = sample_reference_frames(50, 4, axis_fixed=None, device="cpu").numpy()
    torch.manual_seed(99)
    res["mc_randn"] = torch.randn((200, 4)).numpy()
    np.savez_compressed(os.path.join(OUT, "frames.npz"), **res)
    print("frames", res["frames_none"].shape, res["frames_axis2"].shape, res["mc_frames"].shape)


if __name__ == "__main__":
    main()
