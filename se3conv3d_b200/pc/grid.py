import torch

from .._lib import lib, check, ptr, stream, workspace, grid_setup, num_batches, LazyAttrs


class BoundingBox(object):
    """Per-batch axis-aligned bounding box, padded by 1e-6 (pc/BoundingBox.py:10-18)."""

    def __init__(self, p_point_cloud, p_cell_size=1.0):
        self.min_, self.max_, self.num_cells_ = grid_setup(p_point_cloud.pts_, p_point_cloud.batch_ids_,
                                                           num_batches(p_point_cloud), p_cell_size, 1e-6)

    def __repr__(self):
        return "### Min:\n{}\n### Max:\n{}".format(self.min_, self.max_)


class Grid(LazyAttrs):
    """Regular voxel grid over a point cloud (pc/Grid.py:9-58).

    `cell_ids_` [N] are dense cell ranks in sorted-key order, `sorted_ids_` = argsort(cell_ids_),
    `sorted_cell_ids_` the ranks in that order; additionally `num_used_cells_` and `cell_ends_`
    (inclusive segment ends, int32) feed the segment-pooling kernel.  The whole construction is two
    native calls (bounding box + extents; keys -> radix sort -> ranks) and ONE host read (the number
    of occupied cells), instead of the reference's scatter / unique / argsort chain."""

    def __init__(self, p_point_cloud, p_cell_size):
        self.pointcloud_ = p_point_cloud
        self.cell_size_ = p_cell_size
        self.bounding_box_ = BoundingBox(p_point_cloud, p_cell_size)
        self.num_cells_ = self.bounding_box_.num_cells_
        self.cell_ids_ = None
        self.sorted_ids_ = None
        self.sorted_cell_ids_ = None
        self.__compute_cell_ids__()

    def __compute_cell_ids__(self):
        pc = self.pointcloud_
        pts = pc.pts_.to(torch.float32).contiguous()
        b = pc.batch_ids_.to(torch.int32).contiguous()
        n, dev = pts.shape[0], pts.device
        L = lib()
        ws = workspace(L.se3_grid_cells_workspace_bytes(n), dev, 'grid')
        self.cell_ids_ = torch.empty(n, dtype=torch.int64, device=dev)
        self.sorted_ids_ = torch.empty(n, dtype=torch.int64, device=dev)
        ends = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        m = torch.empty(1, dtype=torch.int64, device=dev)
        check(L.se3_grid_cells(ptr(pts), ptr(b), n, ptr(self.bounding_box_.min_), ptr(self.num_cells_),
                               float(self.cell_size_), ptr(ws), ws.numel(), ptr(self.cell_ids_), ptr(self.sorted_ids_),
                               ptr(ends), ptr(m), 0, stream()), "se3_grid_cells")
        self.num_used_cells_ = int(m.item())
        self.cell_ends_ = ends[:self.num_used_cells_]

    @property
    def sorted_cell_ids_(self):
        if self._sorted_cell_ids is None and self.cell_ids_ is not None:
            self._sorted_cell_ids = self.cell_ids_[self.sorted_ids_]
        return self._sorted_cell_ids

    @sorted_cell_ids_.setter
    def sorted_cell_ids_(self, v):
        self._sorted_cell_ids = v

    def __repr__(self):
        return "### Cell size:\n{}\n### Num cells:\n{}\n### Cell Ids:\n{}\n### Sorted Ids:\n{}\n".format(
            self.cell_size_, self.num_cells_, self.cell_ids_, self.sorted_ids_)
