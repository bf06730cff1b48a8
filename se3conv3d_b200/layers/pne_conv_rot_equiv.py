import math

import torch

from .base import IConvLayer, IConvLayerFactory
from ..custom_ops.functions import RotEquivConv, ACT_CODES, WeightLayoutCache
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

    # dims of the geometry vector per relative-rotation encoding: 3 rotated offsets + 6 / 9 / 4
    REL_ROT_DIMS = {"6D": 9, "matrix": 12, "quaternion": 7}

    def __init__(self, p_dims, p_in_features, p_out_features, p_num_basis, p_pne_type):
        super(PNEConvLayerRotEquiv, self).__init__(p_dims, p_in_features, p_out_features)
        self.num_basis_ = p_num_basis
        self.pne_type_ = p_pne_type
        self.aggregation_ = "add"
        if "mlp" not in p_pne_type:
            raise Exception("KPNE convolution not implemeted yet for Rot Equiv.")
        if p_pne_type not in ACT_CODES and p_pne_type != "mlp_softmax":
            raise Se3Error("pne type %r is not supported (supported: %s, mlp_softmax)" % (p_pne_type, sorted(ACT_CODES)))
        if int(p_num_basis) not in (8, 16, 32, 64):
            raise Se3Error("p_num_basis must be 8, 16, 32 or 64 (the aggregation op's basis counts, "
                           "feat_basis_utils.cuh:35-41); the fused kernels need 32, got %r" % (p_num_basis,))
        if int(p_in_features) < 1 or int(p_out_features) < 1:
            raise Se3Error("channel counts must be positive")
        # same construction order as the reference so a given torch seed yields the same init
        stddev = math.sqrt(1.0 / p_dims)
        self.proj_axes_ = torch.nn.Parameter(torch.empty(p_dims, p_num_basis))
        self.proj_axes_.data.uniform_(-stddev, stddev)
        self.proj_biases_ = torch.nn.Parameter(torch.zeros((p_num_basis,), dtype=torch.float32))
        self.conv_weights_ = torch.nn.Parameter(torch.empty(p_in_features, p_num_basis, p_out_features))
        stdv = math.sqrt(1.0 / (p_in_features * p_num_basis))
        self.conv_weights_.data.uniform_(-stdv, stdv)

    def _fused_ok(self, p_pc_in, p_pc_out):
        """The fused kernels cover the shipped configuration space: '6D' encoding, an elementwise activation, 32 basis
        functions, at most 4 frames (and, on the tensor-core path, an output width that is a multiple of 8)."""
        if PNEConvLayerRotEquiv.rel_rot_type != "6D" or self.pne_type_ not in ACT_CODES or int(self.num_basis_) != 32:
            return False
        if self.dims_ != 9 or max(int(getattr(p_pc_in, "n_frames_", 1)), int(getattr(p_pc_out, "n_frames_", 1))) > 4:
            return False
        return int(self.precision) == 0 or self.feat_output_size_ % 8 == 0

    def __compute_convolution__(self, p_pc_in, p_pc_out, p_in_features, p_neighborhood):
        rel = PNEConvLayerRotEquiv.rel_rot_type
        if rel not in self.REL_ROT_DIMS:
            raise ValueError("rel_rot_type must be one of %s" % sorted(self.REL_ROT_DIMS))
        if self.dims_ != self.REL_ROT_DIMS[rel]:
            raise Se3Error("p_dims = %d does not match the %r relative-rotation encoding (%d geometry components)" %
                           (self.dims_, rel, self.REL_ROT_DIMS[rel]))
        if not self._fused_ok(p_pc_in, p_pc_out):
            return self._composed_convolution(p_pc_in, p_pc_out, p_in_features, p_neighborhood)
        geom = p_neighborhood.conv_geometry(p_pc_in, p_pc_out)
        norm_dist, norm_num = self._host_scalars()
        wc = self.__dict__.get("_wcache")
        if wc is None:
            wc = self.__dict__["_wcache"] = WeightLayoutCache()
        return RotEquivConv.apply(p_in_features, self.proj_axes_, self.proj_biases_, self.conv_weights_, geom,
                                  ACT_CODES[self.pne_type_], int(self.precision), norm_dist, norm_num / geom.f_in, wc)

    def _composed_convolution(self, p_pc_in, p_pc_out, p_in_features, p_neighborhood):
        """The configurations outside the fused kernels ('matrix' / 'quaternion' relative rotations,
        pc/RotationFunctions.py:593-600; the `mlp_softmax` basis, layers/PNEConvLayer.py:97; 8 / 16 / 64 basis
        functions; an output width that is not a multiple of 8 at precision 1): the statement sequence of the reference
        (layers/PNEConvLayerRotEquiv.py:61-128, 199-216) on the GPU -- geometry and basis as tensor ops, the aggregation
        through this package's `FeatBasisProj` op (se3_feat_basis_proj / _grad), autograd for the rest.  Not fused,
        fp32; no shipped config takes this branch."""
        from ..custom_ops import FeatBasisProj
        from ..pc.rotation_functions import change_direction_to_local_frame, get_relative_rot
        nb = p_neighborhood.neighbors_.to(torch.int64)
        fo, fi = int(p_pc_out.n_frames_), int(p_pc_in.n_frames_)
        with torch.no_grad():
            i, j = nb[:, 0], nb[:, 1]
            e = i.shape[0]
            rel_pt = (p_pc_in.pts_[j] - p_pc_out.pts_[i]) * self.norm_neigh_dist_
            u = change_direction_to_local_frame(rel_pt, p_pc_out.local_frames_[i])             # [E, fo, 3]
            u = u[:, :, None, :].expand(e, fo, fi, 3).reshape(e, fo * fi, 3)
            r = get_relative_rot(p_pc_out.local_frames_[i], p_pc_in.local_frames_[j], PNEConvLayerRotEquiv.rel_rot_type)
            g = torch.cat((u, r), dim=-1).reshape(e * fo * fi, -1)
            a = torch.arange(fo, device=nb.device)[None, :, None]
            b = torch.arange(fi, device=nb.device)[None, None, :]
            rows = (i[:, None, None] * fo + a).expand(e, fo, fi).reshape(-1)
            cols = (j[:, None, None] * fi + b).expand(e, fo, fi).reshape(-1)
            # edges are grouped by sample already: a stable sort by expanded row keeps the reference's grouping
            order = torch.sort(rows, stable=True)[1]
            g, rows, cols = g[order], rows[order], cols[order]
            m_rows = int(p_pc_out.pts_.shape[0]) * fo
            ends = torch.cumsum(torch.bincount(rows, minlength=m_rows), 0).to(torch.int32)
            nbe = torch.stack((rows, cols), dim=1)
        pre = torch.matmul(g, self.proj_axes_) + self.proj_biases_.reshape(1, -1)
        if self.pne_type_ == "mlp_softmax":
            basis = torch.softmax(pre, dim=-1)
        elif self.pne_type_ == "mlp_gelu":
            basis = torch.nn.functional.gelu(pre)
        elif self.pne_type_ == "mlp_relu":
            basis = torch.relu(pre)
        elif self.pne_type_ == "mlp_sin":
            basis = torch.sin(pre)
        else:
            basis = pre
        t = FeatBasisProj.apply(basis, p_in_features, nbe, ends)
        y = torch.einsum("nik,iko->no", t, self.conv_weights_)
        return y / fi * self.norm_num_neighs_


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
