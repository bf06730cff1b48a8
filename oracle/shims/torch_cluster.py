"""TEST INFRASTRUCTURE ONLY -- brute-force stand-in for torch_cluster 1.6.1 (knn, radius).  See
oracle/shims/torch_scatter.py."""
import torch


def knn(x, y, k, batch_x=None, batch_y=None, **kw):
    """For every y the k nearest x of the same batch item -> [2, M*k] (row 0: y index, row 1: x index)."""
    rows, cols = [], []
    bx = batch_x if batch_x is not None else torch.zeros(x.shape[0], dtype=torch.long)
    by = batch_y if batch_y is not None else torch.zeros(y.shape[0], dtype=torch.long)
    for s in range(0, y.shape[0], 2048):
        d = torch.cdist(y[s:s + 2048].double(), x.double())
        d[by[s:s + 2048, None] != bx[None, :]] = float("inf")
        kk = min(k, x.shape[0])
        dist, idx = torch.topk(d, kk, dim=1, largest=False)
        ok = torch.isfinite(dist)
        r = torch.arange(s, s + d.shape[0])[:, None].expand_as(idx)
        rows.append(r[ok])
        cols.append(idx[ok])
    return torch.stack((torch.cat(rows), torch.cat(cols)))


def knn_graph(x, k, batch=None, loop=False, **kw):
    return knn(x, x, k if loop else k + 1, batch, batch)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, **kw):
    rows, cols = [], []
    bx = batch_x if batch_x is not None else torch.zeros(x.shape[0], dtype=torch.long)
    by = batch_y if batch_y is not None else torch.zeros(y.shape[0], dtype=torch.long)
    for s in range(0, y.shape[0], 2048):
        d = torch.cdist(y[s:s + 2048], x)
        ok = (d < r) & (by[s:s + 2048, None] == bx[None, :])
        rr, cc = torch.nonzero(ok, as_tuple=True)
        rows.append(rr + s)
        cols.append(cc)
    return torch.stack((torch.cat(rows), torch.cat(cols)))


def fps(*a, **kw):
    raise NotImplementedError("fps is not used by any shipped rot config")
