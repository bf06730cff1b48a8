"""TEST INFRASTRUCTURE ONLY -- pure-torch stand-in for torch_scatter 2.1.1 (not installed, no
network) so the *reference* Python package can be imported by tests/golden/gen_*.py and by the
reference arm of bench.py.  Never imported by se3conv3d_b200/."""
import torch


def _n(index, dim_size):
    return int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)


def _ex(index, src):
    return index if src.dim() == 1 else index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    n = _n(index, dim_size)
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index.to(torch.int64), src)


scatter_sum = scatter_add


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    n = _n(index, dim_size)
    s = scatter_add(src, index, 0, dim_size=n)
    ones = torch.ones(index.shape[0], dtype=src.dtype if src.is_floating_point() else torch.float32,
                      device=src.device)
    cnt = torch.zeros(n, dtype=ones.dtype, device=src.device).index_add_(0, index.to(torch.int64), ones).clamp_(min=1)
    cnt = cnt.view(-1, *([1] * (src.dim() - 1)))
    return s / cnt if src.is_floating_point() else torch.div(s, cnt.to(s.dtype), rounding_mode="floor")


def _red(src, index, dim_size, how):
    n = _n(index, dim_size)
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res.scatter_reduce_(0, _ex(index.to(torch.int64), src), src, reduce=how, include_self=False)
    return res, None


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    return _red(src, index, dim_size, "amax")


def scatter_min(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    return _red(src, index, dim_size, "amin")
