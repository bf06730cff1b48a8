import torch

from .._lib import LazyAttrs

from ..scatter import scatter_add, scatter_max, scatter_mean, scatter_min


def pool_by_index(p_in_tensor, p_index, p_pooling_method):
    """avg / max / min / sum pooling of rows by an int64 index (pc/Pointcloud.py:56-74)."""
    if p_pooling_method == "max":
        return scatter_max(p_in_tensor, p_index, dim=0)[0]
    if p_pooling_method == "min":
        return scatter_min(p_in_tensor, p_index, dim=0)[0]
    if p_pooling_method == "avg":
        return scatter_mean(p_in_tensor, p_index, dim=0)
    if p_pooling_method == "sum":
        return scatter_add(p_in_tensor, p_index, dim=0)
    raise ValueError("unknown pooling method " + str(p_pooling_method))


class Pointcloud(LazyAttrs):
    """A batch of point clouds: `pts_` [N,D], `batch_ids_` [N], `batch_size_`
    (same constructor and attributes as pc/Pointcloud.py:6-30)."""

    def __init__(self, p_pts, p_batch_ids, **kwargs):
        self.pts_with_grads_ = bool(kwargs.pop("requires_grad", False))
        self.batch_size_host_ = kwargs.pop("batch_size_host", None)  # optional host copy: avoids a sync
        self.pts_ = torch.as_tensor(p_pts, **kwargs)
        self.batch_ids_ = torch.as_tensor(p_batch_ids, **kwargs)
        self._batch_size = None  # device scalar max(batch_ids_) + 1, built on first use
        if self.pts_with_grads_:
            self.pts_.requires_grad = True

    @property
    def batch_size_(self):
        if self._batch_size is None:
            if self.batch_size_host_ is not None:
                self._batch_size = torch.tensor(self.batch_size_host_, dtype=self.batch_ids_.dtype,
                                                device=self.batch_ids_.device)
            else:
                self._batch_size = torch.max(self.batch_ids_) + 1
        return self._batch_size

    @batch_size_.setter
    def batch_size_(self, v):
        self._batch_size = v

    def to_device(self, p_device):
        self.pts_ = self.pts_.to(p_device)
        self.batch_ids_ = self.batch_ids_.to(p_device)
        self.batch_size_ = self.batch_size_.to(p_device)

    def get_num_points_per_batch(self):
        with torch.no_grad():
            return torch.bincount(self.batch_ids_.to(torch.int64), minlength=int(self.batch_size_)).to(torch.int32)

    def global_pooling(self, p_in_tensor, p_pooling_method="avg"):
        return pool_by_index(p_in_tensor, self.batch_ids_.to(torch.int64), p_pooling_method)

    def global_upsample(self, p_in_tensor):
        return torch.index_select(p_in_tensor, 0, self.batch_ids_.to(torch.int64))

    def __repr__(self):
        return "### Points:\n{}\n### Batch Ids:\n{}\n### Batch Size:\n{}".format(self.pts_, self.batch_ids_,
                                                                               self.batch_size_)
