"""Small pure-torch scatter helpers for HOST-SIDE bookkeeping (bounding boxes, counts, pooling of
labels).  They stand in for the `torch_scatter` calls the reference makes outside the hot path
(pc/BoundingBox.py:17-18, custom_ops/BallQuery.py:36-37, pc/GridSubSample.py:43-45)."""
import torch


def _dim_size(index, dim_size):
    if dim_size is not None:
        return int(dim_size)
    return int(index.max().item()) + 1 if index.numel() else 0


def _expand(index, src):
    if src.dim() == 1:
        return index
    return index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)


def scatter_add(src, index, dim=0, dim_size=None):
    assert dim == 0
    n = _dim_size(index, dim_size)
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index.to(torch.int64), src)


def scatter_mean(src, index, dim=0, dim_size=None):
    assert dim == 0
    n = _dim_size(index, dim_size)
    s = scatter_add(src, index, 0, n)
    cnt = torch.zeros(n, dtype=src.dtype if src.is_floating_point() else torch.float32, device=src.device)
    cnt.index_add_(0, index.to(torch.int64), torch.ones_like(index, dtype=cnt.dtype))
    cnt = cnt.clamp_(min=1)
    if src.is_floating_point():
        return s / cnt.view(-1, *([1] * (src.dim() - 1)))
    return torch.div(s, cnt.view(-1, *([1] * (src.dim() - 1))).to(s.dtype), rounding_mode="floor")


def _scatter_reduce(src, index, n, reduce):
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.scatter_reduce_(0, _expand(index.to(torch.int64), src), src, reduce=reduce, include_self=False)


def scatter_max(src, index, dim=0, dim_size=None):
    assert dim == 0
    return _scatter_reduce(src, index, _dim_size(index, dim_size), "amax"), None


def scatter_min(src, index, dim=0, dim_size=None):
    assert dim == 0
    return _scatter_reduce(src, index, _dim_size(index, dim_size), "amin"), None
