"""Helpers shared by tests/golden/gen_fpn_golden.py (runs the REFERENCE on CPU in the build container) and
tests/test_gpu_fpn_reference_models.py (runs the unmodified reference models/ over this package on the GPU):
a parameter initialisation that depends only on the parameter's NAME and SHAPE, so that both sides hold identical
weights without shipping a 37 MB state_dict -- it also checks the state_dict naming contract, since a renamed
parameter would get different values."""
import hashlib
import math

import torch

FPN_CFG = {
    "init_subsample": 0.04, "grid_subsamples": [0.05, 0.1, 0.2, 0.4],
    "RefFrames": {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": 2},
}
FULL_GRADS = ("SEG_CONV_.proj_axes_", "ENCODER_.PATCH_EMB_.CONV_LAYERS_.0.conv_weights_", "ENCODER_.CONV_DOWN_.0.conv_weights_",
              "ENCODER_.BLOCKS_LIST_.1.0.spatial_conv_.proj_biases_", "SEG_LINEAR_.weight")


def _seed(name):
    return int.from_bytes(hashlib.sha256(name.encode()).digest()[:4], "little")


def reinit_by_name(model):
    """Deterministic values per parameter name: conv weights / axes / Linear weights uniform with the fan-in scale,
    biases small, skip gammas in [0.5, 1.5] (the reference's 1e-6 init would hide every convolution behind its skip),
    BatchNorm affine near identity."""
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            g = torch.Generator().manual_seed(_seed(name))
            u = torch.rand(p.shape, generator=g, dtype=torch.float64)
            if name.endswith("gamma_"):
                v = 0.5 + u
            elif name.endswith("conv_weights_"):
                v = (2 * u - 1) * math.sqrt(3.0 / (p.shape[0] * p.shape[1]))
            elif name.endswith("proj_axes_"):
                v = (2 * u - 1) * math.sqrt(1.0 / p.shape[0])
            elif name.endswith("layer_.weight"):          # BatchNorm scale
                v = 0.8 + 0.4 * u
            elif p.dim() == 2:                             # Linear weight [out, in]
                v = (2 * u - 1) * math.sqrt(3.0 / p.shape[1])
            else:                                          # biases (conv basis, Linear, BatchNorm shift)
                v = 0.2 * (2 * u - 1)
            p.copy_(v.to(p.dtype))
