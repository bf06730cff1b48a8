"""The block around the convolution: BatchNormPC, DropPathPC, SkipConnection, Block and
ResNetFormer with the reference's constructors, attribute / parameter names (so checkpoints load) and forward
signatures (layers/BatchNormPC.py:7-31, DropPathPC.py:5-49, SkipConnection.py:7-42, Block.py:5-50,
ResNetFormer.py:5-90).  The convolution inside is the fused kernel; the gamma-skip with its drop path is one kernel each way
(csrc/block_ops.cu); normalisation and the two Linear layers are library ops (cuDNN / cuBLAS through torch)."""
import torch

from .base import PreProcessModule


class NormLayerPC(torch.nn.Module):
    """Normalisation layer interface: forward(x, point cloud) (layers/NormLayerPC.py)."""

    def __init__(self, p_num_features):
        super(NormLayerPC, self).__init__()
        self.num_features_ = p_num_features


class BatchNormPC(NormLayerPC):
    """BatchNorm1d over all (point, frame) rows, momentum 0.2."""

    def __init__(self, p_num_features):
        super(BatchNormPC, self).__init__(p_num_features)
        self.layer_ = torch.nn.BatchNorm1d(p_num_features, momentum=0.2)

    def forward(self, p_x, p_pc):
        return self.layer_(p_x)


class DropPathPC(torch.nn.Module):
    """Stochastic depth per batch item: all rows of a dropped item are zeroed, kept items are rescaled."""

    def __init__(self, p_drop_prob):
        super(DropPathPC, self).__init__()
        self.drop_prob_ = p_drop_prob

    def forward(self, p_x, p_pc):
        if self.drop_prob_ == 0.0 or not self.training:
            return p_x
        keep = 1.0 - self.drop_prob_
        n_items = int(getattr(p_pc, "batch_size_host_", None) or p_pc.batch_size_)
        mask = torch.floor(keep + torch.rand((n_items,), dtype=p_x.dtype, device=p_x.device))
        ids = p_pc.batch_ids_considering_frames_ if hasattr(p_pc, "batch_ids_considering_frames_") else p_pc.batch_ids_
        return p_x / keep * mask[ids.to(torch.int64)].reshape(-1, 1)


class SkipConnection(torch.nn.Module):
    """out = drop_path(x * gamma) + y with a learnable per-channel gamma initialised to 1e-6
    (layers/SkipConnection.py:7-42).  On CUDA tensors the scale, the per-item drop-path mask, its 1 / keep rescale and
    the skip add are ONE kernel forward (se3_gamma_skip_fwd) and one backward (+ an ordered reduction for d gamma)."""

    def __init__(self, p_drop_prob, p_num_features, p_init_gamma=1e-6):
        super(SkipConnection, self).__init__()
        self.drop_path_ = DropPathPC(p_drop_prob)
        self.gamma_ = torch.nn.Parameter(p_init_gamma * torch.ones((1, p_num_features)))

    def forward(self, p_x, p_y, p_pc):
        if not (p_x.is_cuda and p_x.dtype == torch.float32 and p_y.dtype == torch.float32 and p_x.dim() == 2):
            return self.drop_path_(p_x * self.gamma_, p_pc) + p_y
        from ..custom_ops.functions import GammaSkip, _point_items
        n_points = int(p_pc.batch_ids_.shape[0])
        frames = max(int(p_x.shape[0]) // max(n_points, 1), 1)
        scale, items = None, None
        p = self.drop_path_.drop_prob_
        if p != 0.0 and self.training:
            keep = 1.0 - p
            n_items = int(getattr(p_pc, "batch_size_host_", None) or p_pc.batch_size_)
            # the same draw as DropPathPC: floor(keep + U[0,1)) per batch item, divided by keep
            scale = torch.floor(keep + torch.rand((n_items,), dtype=torch.float32, device=p_x.device)) / keep
            items = _point_items(p_pc)
        return GammaSkip.apply(p_x, p_y, self.gamma_, scale, items, frames)


class Block(PreProcessModule):
    """Block interface: forward(point cloud, features, neighbourhood)."""

    def __init__(self, p_in_features, p_out_features, p_conv_fact, p_norm_layer, p_path_drop_prob):
        super(Block, self).__init__()
        self.feat_input_size_ = p_in_features
        self.feat_output_size_ = p_out_features


class ResNetFormer(Block):
    """norm -> spatial conv -> gamma-skip -> norm -> Linear(x2) -> GELU -> Linear -> gamma-skip."""

    def __init__(self, p_in_features, p_out_features, p_conv_fact, p_norm_layer, p_path_drop_prob):
        super(ResNetFormer, self).__init__(p_in_features, p_out_features, p_conv_fact, p_norm_layer, p_path_drop_prob)
        self.act_func_ = torch.nn.GELU()
        self.feat_scale_factor_ = 2
        cin, cout = self.feat_input_size_, self.feat_output_size_
        self.spatial_conv_ = p_conv_fact.create_conv_layer(cin, cin)
        self.norm_1_ = p_norm_layer(cin)
        self.norm_2_ = p_norm_layer(cin)
        self.linear_1_ = torch.nn.Linear(cin, cin * self.feat_scale_factor_)
        self.linear_2_ = torch.nn.Linear(cin * self.feat_scale_factor_, cout)
        self.skip_path_1_ = SkipConnection(p_path_drop_prob, cin)
        self.skip_path_2_ = SkipConnection(p_path_drop_prob, cout)
        if cin != cout:
            self.skip_conv_ = torch.nn.Linear(cin, cout)

    def forward(self, p_pc_in, p_in_features, p_neighborhood):
        x = self.norm_1_(p_in_features, p_pc_in)
        x = self.spatial_conv_(p_pc_in=p_pc_in, p_pc_out=p_pc_in, p_in_features=x, p_neighborhood=p_neighborhood)
        x = self.skip_path_1_(x, p_in_features, p_pc_in)
        y = self.linear_2_(self.act_func_(self.linear_1_(self.norm_2_(x, p_pc_in))))
        skip = self.skip_conv_(x) if self.feat_input_size_ != self.feat_output_size_ else x
        return self.skip_path_2_(y, skip, p_pc_in)
