"""Generates tests/golden/variant_*.npz: the reference's own PNEConvLayerRotEquiv (forward + autograd backward, float64,
CPU) for the configurations outside the fused kernels -- 'matrix' and 'quaternion' relative-rotation encodings
(pc/RotationFunctions.py:593-600, selected through PNEConvLayerRotEquiv.rel_rot_type, layers/PNEConvLayerRotEquiv.py:54),
the `mlp_softmax` basis (layers/PNEConvLayer.py:97) and 16 basis functions.  Same substitutions as gen_layer_golden.py
(FeatBasisProj -> scatter formulation).

    python tests/golden/gen_variant_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.ref_import import import_reference  # noqa: E402
from gen_layer_golden import brute_radius  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (rel_rot, p_dims, pne, num_basis, N, F, Cin, Cout, radius)
    "matrix": ("matrix", 12, "mlp_gelu", 32, 80, 2, 8, 16, 0.45),
    "quaternion": ("quaternion", 7, "mlp_gelu", 32, 80, 2, 8, 16, 0.45),
    "softmax": ("6D", 9, "mlp_softmax", 32, 80, 2, 8, 16, 0.45),
    "basis16": ("6D", 9, "mlp_gelu", 16, 80, 1, 8, 24, 0.45),
}


def main():
    import_reference()
    refmod = sys.modules["point_cloud_lib.layers.PNEConvLayerRotEquiv"]
    from point_cloud_lib.pc.RotationFunctions import quaternion_to_matrix

    class ScatterFeatBasisProj:
        @staticmethod
        def apply(basis, feats, neighbors, ends):
            t = torch.zeros((ends.shape[0], feats.shape[1], basis.shape[1]), dtype=feats.dtype)
            return t.index_add(0, neighbors[:, 0].long(), feats[neighbors[:, 1].long()][:, :, None] * basis[:, None, :])
    refmod.FeatBasisProj = ScatterFeatBasisProj
    for name, (rel, dims, pne, nbasis, n, f, cin, cout, radius) in CASES.items():
        gen = torch.Generator().manual_seed(sum(map(ord, name)))
        pts = torch.rand(n, 3, generator=gen)
        b = torch.sort(torch.randint(0, 2, (n,), generator=gen))[0]
        q = torch.randn(n * f, 4, generator=gen)
        fr = quaternion_to_matrix(q / q.norm(dim=1, keepdim=True)).reshape(n, f, 9)
        nb, ends = brute_radius(pts, pts, b, b, radius)
        x = torch.randn(n * f, cin, generator=gen)
        dy = torch.randn(n * f, cout, generator=gen)
        refmod.PNEConvLayerRotEquiv.rel_rot_type = rel
        refmod.PNEConvLayerRotEquiv.empty_rot_tenors_cache()
        torch.manual_seed(4321)
        lay = refmod.PNEConvLayerRotEquiv(dims, cin, cout, nbasis, pne).double()
        with torch.no_grad():
            lay.proj_biases_.copy_(0.1 * torch.randn(nbasis, generator=gen))
        lay.norm_neigh_dist_ = torch.tensor(1.0 / radius, dtype=torch.float64)
        lay.norm_num_neighs_ = torch.tensor(n / nb.shape[0], dtype=torch.float64)
        pc = types.SimpleNamespace(pts_=pts.double(), local_frames_=fr.double(), n_frames_=f)
        neigh = types.SimpleNamespace(neighbors_=nb.clone(), start_ids_=ends.clone())
        xr = x.double().clone().requires_grad_(True)
        y = lay(pc, pc, xr, neigh)
        (y * dy.double()).sum().backward()
        np.savez_compressed(
            os.path.join(OUT, "variant_%s.npz" % name), pts=pts.numpy(), frames=fr.numpy(), neighbors=nb.numpy(),
            ends=ends.numpy(), x=x.numpy(), dy=dy.numpy(), proj_axes=lay.proj_axes_.detach().float().numpy(),
            proj_biases=lay.proj_biases_.detach().float().numpy(), conv_weights=lay.conv_weights_.detach().float().numpy(),
            norm_neigh_dist=np.float32(1.0 / radius), norm_num_neighs=np.float32(n / nb.shape[0]), rel=np.array(rel),
            pne=np.array(pne), dims=np.int32(dims), y=y.detach().numpy(), dx=xr.grad.numpy(),
            dW=lay.conv_weights_.grad.numpy(), dA=lay.proj_axes_.grad.numpy(), dB=lay.proj_biases_.grad.numpy())
        print(name, "E =", nb.shape[0], "y", tuple(y.shape))
    refmod.PNEConvLayerRotEquiv.rel_rot_type = "6D"


if __name__ == "__main__":
    main()
