from abc import ABC, abstractmethod

import torch

from ..pc import KnnNeighborhood, BQNeighborhood


class PreProcessModule(torch.nn.Module):
    """Module with a recursive pre-process switch (layers/PreProcessModule.py:3-52)."""

    def __init__(self):
        self.pre_process_ = False
        super(PreProcessModule, self).__init__()

    def _set_children(self, p_module, p_start):
        for child in p_module.children():
            if isinstance(child, PreProcessModule):
                child.start_pre_process() if p_start else child.end_pre_process()
            elif isinstance(child, torch.nn.ModuleList):
                self._set_children(child, p_start)

    def start_pre_process(self):
        self.pre_process_ = True
        self._set_children(self, True)

    def end_pre_process(self):
        self.pre_process_ = False
        self._set_children(self, False)


class IConvLayer(PreProcessModule, ABC):
    """Point-convolution interface: buffers `norm_neigh_dist_` / `norm_num_neighs_` (EMA 0.9/0.1,
    updated only while pre-processing) and the four-argument forward (layers/IConvLayer.py:8-104)."""

    def __init__(self, p_dims, p_in_features, p_out_features):
        super(IConvLayer, self).__init__()
        self.dims_ = p_dims
        self.feat_input_size_ = p_in_features
        self.feat_output_size_ = p_out_features
        self.register_buffer("norm_neigh_dist_", torch.tensor(0, dtype=torch.float32))
        self.register_buffer("norm_num_neighs_", torch.tensor(0, dtype=torch.float32))

    @abstractmethod
    def __compute_convolution__(self, p_pc_in, p_pc_out, p_in_features, p_neighborhood):
        pass

    def _host_scalars(self):
        """Host copies of the two scalar buffers, refreshed only when a buffer was replaced or written
        (pre-processing, load_state_dict, .fill_) -- so a steady-state forward never synchronises."""
        key = (id(self.norm_neigh_dist_), self.norm_neigh_dist_._version, id(self.norm_num_neighs_),
               self.norm_num_neighs_._version)
        cache = getattr(self, "_scalar_cache", None)
        if cache is None or cache[0] != key:
            # (the buffers are held by the entry: their ids cannot be recycled while it is cached)
            cache = (key, float(self.norm_neigh_dist_), float(self.norm_num_neighs_), self.norm_neigh_dist_,
                     self.norm_num_neighs_)
            self._scalar_cache = cache
        return cache[1], cache[2]

    def forward(self, p_pc_in, p_pc_out, p_in_features, p_neighborhood):
        if self.pre_process_:
            with torch.no_grad():
                if isinstance(p_neighborhood, BQNeighborhood):
                    new_dist = torch.tensor(1.0 / p_neighborhood.radius_, dtype=torch.float32)
                elif isinstance(p_neighborhood, KnnNeighborhood):
                    diff = p_pc_in.pts_[p_neighborhood.neighbors_[:, 1].long(), :] - \
                        p_pc_out.pts_[p_neighborhood.neighbors_[:, 0].long(), :]
                    mean_dist = torch.mean(torch.sqrt(torch.sum(diff ** 2, -1))).item()
                    new_dist = torch.tensor(1.0 / (2.0 * mean_dist), dtype=torch.float32)
                else:
                    raise TypeError("unknown neighbourhood type")
                dev = self.norm_neigh_dist_.device
                self.norm_neigh_dist_ = 0.9 * self.norm_neigh_dist_ + 0.1 * new_dist.to(dev)
                new_num = torch.tensor(p_neighborhood.start_ids_.shape[0] / p_neighborhood.neighbors_.shape[0],
                                       dtype=torch.float32)
                self.norm_num_neighs_ = 0.9 * self.norm_num_neighs_ + 0.1 * new_num.to(dev)
        return self.__compute_convolution__(p_pc_in, p_pc_out, p_in_features, p_neighborhood)


class IConvLayerFactory(ABC):
    """Layer factory interface (layers/IConvLayer.py:107-159)."""

    def __init__(self, p_dims):
        super(IConvLayerFactory, self).__init__()
        self.dims_ = p_dims
        self.conv_list_ = []

    def update_parameters(self, **kwargs):
        pass

    @abstractmethod
    def __create_conv_layer_imp__(self, p_in_features, p_out_features):
        pass

    def create_conv_layer(self, p_in_features, p_out_features):
        conv = self.__create_conv_layer_imp__(p_in_features, p_out_features)
        self.conv_list_.append(conv)
        return conv
