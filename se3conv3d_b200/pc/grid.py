import torch

from ..custom_ops import ComputeKeys
from ..scatter import scatter_max, scatter_min


class BoundingBox(object):
    """Per-batch axis-aligned bounding box, padded by 1e-6 (pc/BoundingBox.py:10-18)."""

    def __init__(self, p_point_cloud):
        idx = p_point_cloud.batch_ids_.to(torch.int64)
        self.max_ = scatter_max(p_point_cloud.pts_, idx, dim=0)[0] + 1e-6
        self.min_ = scatter_min(p_point_cloud.pts_, idx, dim=0)[0] - 1e-6

    def __repr__(self):
        return "### Min:\n{}\n### Max:\n{}".format(self.min_, self.max_)


class Grid(object):
    """Regular voxel grid over a point cloud (pc/Grid.py:9-58).

    `cell_ids_` [N] are dense cell ranks in sorted-key order, `sorted_ids_` = argsort(cell_ids_),
    `sorted_cell_ids_` the ranks in that order; additionally `num_used_cells_` and `cell_ends_`
    (inclusive segment ends, int32) feed the segment-pooling kernel."""

    def __init__(self, p_point_cloud, p_cell_size):
        self.pointcloud_ = p_point_cloud
        self.bounding_box_ = BoundingBox(p_point_cloud)
        self.cell_size_ = p_cell_size
        extent = (self.bounding_box_.max_ - self.bounding_box_.min_) / self.cell_size_
        self.num_cells_ = torch.max(extent.to(torch.int32) + 1, dim=0)[0]
        self.cell_ids_ = None
        self.sorted_ids_ = None
        self.sorted_cell_ids_ = None
        self.__compute_cell_ids__()

    def __compute_cell_ids__(self):
        pc = self.pointcloud_
        cell = torch.full((self.num_cells_.shape[0],), float(self.cell_size_), dtype=torch.float32,
                          device=self.num_cells_.device)
        keys = ComputeKeys.apply(pc.pts_, pc.batch_ids_, self.bounding_box_.min_, self.num_cells_, cell)
        uniq, self.cell_ids_ = torch.unique(keys, return_inverse=True)
        self.num_used_cells_ = int(uniq.shape[0])
        self.sorted_ids_ = torch.argsort(self.cell_ids_, stable=True)
        self.sorted_cell_ids_ = self.cell_ids_[self.sorted_ids_]
        counts = torch.bincount(self.cell_ids_, minlength=self.num_used_cells_)
        self.cell_ends_ = torch.cumsum(counts, 0).to(torch.int32)

    def __repr__(self):
        return "### Cell size:\n{}\n### Num cells:\n{}\n### Cell Ids:\n{}\n### Sorted Ids:\n{}\n".format(
            self.cell_size_, self.num_cells_, self.cell_ids_, self.sorted_ids_)
