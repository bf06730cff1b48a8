from abc import ABC, abstractmethod

import torch

from .._lib import lib, check, ptr, stream
from ..scatter import scatter_max
from .grid import Grid


class SubSample(ABC):
    """Sub-sampling interface (pc/SubSample.py:8-56)."""

    def __init__(self, p_pc_src):
        self.pc_src_ = p_pc_src
        self.ids_ = None
        self.__compute_subsample__()

    @abstractmethod
    def __compute_subsample__(self):
        pass

    @abstractmethod
    def __subsample_tensor__(self, p_tensor, p_method="avg"):
        pass

    @abstractmethod
    def __upsample_tensor__(self, p_tensor):
        pass

    def __repr__(self):
        return "### Ids:\n{}\n".format(self.ids_)


def _segment_pool_raw(x, grid, mode):
    m = grid.num_used_cells_
    out = torch.empty((m, x.shape[1]), dtype=torch.float32, device=x.device)
    check(lib().se3_segment_pool_f32(ptr(x), x.shape[0], x.shape[1], ptr(grid.sorted_ids_), ptr(grid.cell_ends_), m,
                                     mode, ptr(out), stream()), "se3_segment_pool_f32")
    return out


class _SegmentPool(torch.autograd.Function):
    """Differentiable cell pooling (the reference pools features with torch_scatter's scatter_mean / scatter_max,
    pc/GridSubSample.py:69-72, which are differentiable): mean -> grad[cell] / count; max -> the gradient goes to one
    arg-max row per (cell, channel), the first in point order."""

    @staticmethod
    def forward(ctx, x, grid, mode):
        out = _segment_pool_raw(x, grid, mode)
        ctx.grid, ctx.mode = grid, mode
        if mode == 1:
            ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, grad):
        grid, ids = ctx.grid, ctx.grid.cell_ids_
        if ctx.mode == 0:
            ends = grid.cell_ends_.to(torch.int64)
            counts = torch.diff(ends, prepend=ends.new_zeros(1)).to(grad.dtype).clamp_min(1)
            return (grad / counts[:, None])[ids], None, None
        x, out = ctx.saved_tensors
        n, c = x.shape
        row = torch.arange(n, device=x.device, dtype=torch.int64)[:, None].expand(n, c)
        cand = torch.where(x == out[ids], row, torch.full_like(row, n))
        win = torch.full((out.shape[0], c), n, dtype=torch.int64, device=x.device)
        win.scatter_reduce_(0, ids[:, None].expand(n, c), cand, reduce="amin", include_self=True)
        gx = torch.zeros((n + 1, c), dtype=grad.dtype, device=x.device)
        gx.scatter_(0, win, grad)
        return gx[:n], None, None


def segment_pool(p_tensor, grid, mode):
    """mean (mode 0) / max (mode 1) of float rows per grid cell through se3_segment_pool_f32 (differentiable)."""
    squeeze = p_tensor.dim() == 1
    x = p_tensor.reshape(p_tensor.shape[0], -1).to(torch.float32).contiguous()
    out = _SegmentPool.apply(x, grid, mode) if x.requires_grad else _segment_pool_raw(x, grid, mode)
    return out[:, 0] if squeeze else out


class GridSubSample(SubSample):
    """Voxel-grid sub-sampling: average / max pooling per occupied cell, or one random point per
    cell (pc/GridSubSample.py:11-93)."""

    def __init__(self, p_pc_src, p_cell_size, p_rnd_sample=False):
        self.cell_size_ = p_cell_size
        self.rnd_sample_ = p_rnd_sample
        super(GridSubSample, self).__init__(p_pc_src)

    def __compute_subsample__(self):
        self.grid_ = Grid(self.pc_src_, self.cell_size_)
        if self.rnd_sample_:
            ends = self.grid_.cell_ends_.to(torch.int64)
            counts = torch.diff(ends, prepend=ends.new_zeros(1))
            starts = ends - counts
            pick = torch.rand(counts.shape[0]).to(counts.device) * counts
            self.ids_ = torch.floor(pick).to(torch.int32) + starts.to(torch.int32)

    def __subsample_tensor__(self, p_tensor, p_method="avg"):
        if self.rnd_sample_:
            return p_tensor[self.grid_.sorted_ids_[self.ids_.to(torch.int64)]]
        if p_method == "avg":
            if p_tensor.is_floating_point():
                return segment_pool(p_tensor, self.grid_, 0).to(p_tensor.dtype)
            raise TypeError("avg pooling needs a floating point tensor")
        if p_method == "max":
            if p_tensor.is_floating_point():
                return segment_pool(p_tensor, self.grid_, 1).to(p_tensor.dtype)
            return scatter_max(p_tensor, self.grid_.cell_ids_, dim=0, dim_size=self.grid_.num_used_cells_)[0]
        raise ValueError("unknown pooling method " + str(p_method))

    def __upsample_tensor__(self, p_tensor):
        if self.rnd_sample_:
            target = self.grid_.sorted_ids_[self.ids_.to(torch.int64)]
            out = torch.zeros((self.grid_.cell_ids_.shape[0], p_tensor.shape[-1]), dtype=p_tensor.dtype,
                              device=p_tensor.device)
            out[target] = p_tensor
            return out
        return p_tensor[self.grid_.cell_ids_]
