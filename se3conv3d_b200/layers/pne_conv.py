import math

import torch

from .base import IConvLayer, IConvLayerFactory
from ..custom_ops.functions import RotEquivConv, ACT_CODES
from .._lib import Se3Error
from ..pc.neighborhood import ConvGeometry


class _PlainCloud(object):
    """A Pointcloud seen as a one-frame cloud whose frame is the identity: with it the fused SE(3) kernel
    evaluates exactly the translation-only embedding of the standard layer."""

    def __init__(self, pc):
        self.pts_ = pc.pts_
        n = pc.pts_.shape[0]
        eye = torch.eye(3, dtype=torch.float32, device=pc.pts_.device).reshape(1, 1, 9)
        self.local_frames_ = eye.expand(n, 1, 9).contiguous()
        self.n_frames_ = 1


class PNEConvLayer(IConvLayer):
    """The standard (non-equivariant) point convolution with MLP point-neighbourhood embeddings and "add"
    aggregation -- same constructor, parameters (`proj_axes_` [3,K], `proj_biases_` [K], `conv_weights_`
    [Cin,K,Cout]) and forward as layers/PNEConvLayer.py:52-229 -- on the fused SE(3) kernels.

    With one identity frame per point the kernel's 9-vector is g = [d ; 1,0,0 ; 0,1,0] (d = (p_in - p_out) *
    norm_neigh_dist_), so zero-padding proj_axes_ from [3,K] to [9,K] gives pre = d . A + b, the LinearPNE of
    custom_ops/PNE.py:30-40.  The kernel-point embeddings (`kp_*`) and the "max" aggregation of the reference
    are not on this path and raise."""

    precision = 0

    def __init__(self, p_dims, p_in_features, p_out_features, p_num_basis, p_pne_type, p_aggregation="add"):
        super(PNEConvLayer, self).__init__(p_dims, p_in_features, p_out_features)
        if p_dims != 3:
            raise Se3Error("PNEConvLayer: only 3-D point clouds are supported")
        if p_pne_type not in ACT_CODES:
            raise Se3Error("pne type %r is not supported by the fused kernel (supported: %s)" %
                           (p_pne_type, sorted(ACT_CODES)))
        if p_aggregation != "add":
            raise Se3Error("PNEConvLayer: only the 'add' aggregation runs on the fused kernel")
        self.num_basis_ = p_num_basis
        self.pne_type_ = p_pne_type
        self.aggregation_ = p_aggregation
        stddev = math.sqrt(1.0 / p_dims)
        self.proj_axes_ = torch.nn.Parameter(torch.empty(p_dims, p_num_basis))
        self.proj_axes_.data.uniform_(-stddev, stddev)
        self.proj_biases_ = torch.nn.Parameter(torch.zeros((p_num_basis,), dtype=torch.float32))
        self.conv_weights_ = torch.nn.Parameter(torch.empty(p_in_features, p_num_basis, p_out_features))
        stdv = math.sqrt(1.0 / (p_in_features * p_num_basis))
        self.conv_weights_.data.uniform_(-stdv, stdv)

    @staticmethod
    def _plain(pc):
        cached = getattr(pc, "_se3_plain", None)
        if cached is None or cached.pts_ is not pc.pts_:
            cached = _PlainCloud(pc)
            try:
                pc._se3_plain = cached
            except AttributeError:
                pass
        return cached

    def __compute_convolution__(self, p_pc_in, p_pc_out, p_in_features, p_neighborhood):
        pin = self._plain(p_pc_in)
        pout = pin if p_pc_out is p_pc_in else self._plain(p_pc_out)
        geom = p_neighborhood.conv_geometry(pin, pout)
        axes9 = torch.cat((self.proj_axes_, self.proj_axes_.new_zeros(6, self.num_basis_)), dim=0)
        norm_dist, norm_num = self._host_scalars()
        return RotEquivConv.apply(p_in_features, axes9, self.proj_biases_, self.conv_weights_, geom,
                                  ACT_CODES[self.pne_type_], int(self.precision), norm_dist, norm_num)


class PNEConvLayerFactory(IConvLayerFactory):
    """Factory with the reference's signature (layers/PNEConvLayer.py:232-270)."""

    def __init__(self, p_dims, p_num_basis, p_pne_type, p_aggregation="add"):
        super(PNEConvLayerFactory, self).__init__(p_dims)
        self.num_basis_ = p_num_basis
        self.pne_type_ = p_pne_type
        self.aggregation_ = p_aggregation

    def update_parameters(self, **kwargs):
        if "num_basis" in kwargs:
            self.num_basis_ = kwargs["num_basis"]

    def __create_conv_layer_imp__(self, p_in_features, p_out_features):
        return PNEConvLayer(self.dims_, p_in_features, p_out_features, self.num_basis_, self.pne_type_, self.aggregation_)
