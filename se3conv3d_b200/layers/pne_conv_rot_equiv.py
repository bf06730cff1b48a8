import math

import torch

from .base import IConvLayer, IConvLayerFactory
from ..custom_ops.functions import RotEquivConv, ACT_CODES
from .._lib import Se3Error


class PNEConvLayerRotEquiv(IConvLayer):
    """Continuous SE(3) group convolution with point-neighbourhood embeddings.

    Same constructor, parameters (`proj_axes_` [9,K], `proj_biases_` [K], `conv_weights_`
    [Cin,K,Cout]), buffers and forward signature as layers/PNEConvLayerRotEquiv.py:49-216 (+ its
    base layers/PNEConvLayer.py:52-158), so reference checkpoints load unchanged.  The forward is a
    single fused call (se3_conv_fwd) instead of get_rot_tenors -> matmul -> GELU -> FeatBasisProj ->
    einsum; the relative geometry and kernel weights never reach HBM.

    Class attributes:
      rot_tensor_cache / empty_rot_tenors_cache(): kept because the models call it every forward
        (tasks/SemSeg/seg_models.py:92,99,106); geometry is cached on the neighbourhood instead.
      rel_rot_type: '6D' (the only encoding on the B200 path).
      precision: 0 = fp32 CUDA cores (exactness mode), 1 = bf16 tensor cores.
    """

    rot_tensor_cache = {}
    rel_rot_type = "6D"
    precision = 0

    @staticmethod
    def empty_rot_tenors_cache():
        PNEConvLayerRotEquiv.rot_tensor_cache = {}

    def __init__(self, p_dims, p_in_features, p_out_features, p_num_basis, p_pne_type):
        super(PNEConvLayerRotEquiv, self).__init__(p_dims, p_in_features, p_out_features)
        self.num_basis_ = p_num_basis
        self.pne_type_ = p_pne_type
        self.aggregation_ = "add"
        if "mlp" not in p_pne_type:
            raise Exception("KPNE convolution not implemeted yet for Rot Equiv.")
        if p_pne_type not in ACT_CODES:
            raise Se3Error("pne type %r is not supported by the fused kernel (supported: %s)" %
                           (p_pne_type, sorted(ACT_CODES)))
        # same construction order as the reference so a given torch seed yields the same init
        stddev = math.sqrt(1.0 / p_dims)
        self.proj_axes_ = torch.nn.Parameter(torch.empty(p_dims, p_num_basis))
        self.proj_axes_.data.uniform_(-stddev, stddev)
        self.proj_biases_ = torch.nn.Parameter(torch.zeros((p_num_basis,), dtype=torch.float32))
        self.conv_weights_ = torch.nn.Parameter(torch.empty(p_in_features, p_num_basis, p_out_features))
        stdv = math.sqrt(1.0 / (p_in_features * p_num_basis))
        self.conv_weights_.data.uniform_(-stdv, stdv)

    def __compute_convolution__(self, p_pc_in, p_pc_out, p_in_features, p_neighborhood):
        if PNEConvLayerRotEquiv.rel_rot_type != "6D":
            raise Se3Error("only the '6D' relative-rotation encoding is implemented on the B200 path")
        geom = p_neighborhood.conv_geometry(p_pc_in, p_pc_out)
        norm_dist, norm_num = self._host_scalars()
        return RotEquivConv.apply(p_in_features, self.proj_axes_, self.proj_biases_, self.conv_weights_, geom,
                                  ACT_CODES[self.pne_type_], int(self.precision), norm_dist, norm_num / geom.f_in)

class PNEConvLayerRotEquivFactory(IConvLayerFactory):
    """Factory with the reference's signature (layers/PNEConvLayerRotEquiv.py:236-281)."""

    def __init__(self, p_dims, p_num_basis, p_pne_type, p_rel_rot="6D"):
        super(PNEConvLayerRotEquivFactory, self).__init__(p_dims)
        self.num_basis_ = p_num_basis
        self.pne_type_ = p_pne_type
        self.rel_rot_ = p_rel_rot

    def update_parameters(self, **kwargs):
        if "num_basis" in kwargs:
            self.num_basis_ = kwargs["num_basis"]

    def __create_conv_layer_imp__(self, p_in_features, p_out_features):
        PNEConvLayerRotEquiv.rel_rot_type = self.rel_rot_
        return PNEConvLayerRotEquiv(self.dims_, p_in_features, p_out_features, self.num_basis_, self.pne_type_)
