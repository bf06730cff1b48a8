"""Timing of BASELINE configs 3 and 4 on the hot path: fused hierarchy build + one convolution forward / backward per
neighbourhood (bf16 mode), CUDA events, L2 flushed between iterations.

    python tools/config34_run.py 3      # 128 clouds x 1024 points on a sphere, F = 4 sampled frames
    python tools/config34_run.py 4      # one 150k-point room-like scene, F = 1 PCA frames about the up axis
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200.layers import PNEConvLayerRotEquiv  # noqa: E402
from se3conv3d_b200.pc import build_point_hierarchy  # noqa: E402

which = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if which == 3:
    b_items, n_pts, F = 128, 1024, 4
    p = torch.randn(b_items * n_pts, 3, generator=g)
    p = p / p.norm(dim=1, keepdim=True) + 0.01 * torch.randn(b_items * n_pts, 3, generator=g)
    batch = torch.arange(b_items).repeat_interleave(n_pts).to(torch.int32)
    cfg = {"pca": False, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": F}
    init, grids = 0.05, [0.05, 0.1, 0.2, 0.3, 0.4]
    chans = [32, 64, 128, 256, 512, 512]
else:
    b_items, n, F = 1, 150000, 1
    u = torch.rand(n, 3, generator=g)
    sel = torch.randint(0, 4, (n,), generator=g)
    p = torch.stack((u[:, 0] * 8, u[:, 1] * 6, u[:, 2] * 3), 1)
    p[sel == 0, 2] = 0.0
    p[sel == 1, 0] = 0.0
    p[sel == 2, 1] = 0.0
    p[sel == 3] = p[sel == 3] * torch.tensor([0.15, 0.2, 0.3]) + torch.tensor([3.0, 2.0, 0.0])
    p = p + 0.004 * torch.randn(n, 3, generator=g)
    batch = torch.zeros(n, dtype=torch.int32)
    cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": 2, "n_frames": F}
    init, grids = 0.1, [0.2, 0.4, 0.8, 1.6]
    chans = [32, 64, 128, 256, 512]
p, batch = p.to(torch.float32).to(dev), batch.to(dev)
cells = [init] + grids
wanted = []
for l in range(len(cells)):
    wanted.append((l, l, 2.0 * cells[l]))
    if l + 1 < len(cells):
        wanted.append((l, l + 1, 2.0 * cells[l]))
for _ in range(12):
    h, _ = build_point_hierarchy(p, batch, cfg, init, grids, neighborhoods=wanted, n_batches=b_items)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for it in range(10):
    flush.fill_(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h, _ = build_point_hierarchy(p, batch, cfg, init, grids, neighborhoods=wanted, n_batches=b_items)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
sizes = [int(pc.pts_.shape[0]) for pc in h.pcs_]
out = {"config": which, "raw_points": int(p.shape[0]), "frames": F, "level_points": sizes,
       "hierarchy_build_ms_median": round(ts[len(ts) // 2], 3), "layers": []}
tot = 0.0
for nb, (s, t, r) in zip(h.fused_neighborhoods_, wanted):
    cin, cout = chans[s], chans[t]
    pin, pout = h.pcs_[s], h.pcs_[t]
    e = nb.n_edges_
    layer = PNEConvLayerRotEquiv(9, cin, cout, 32, "mlp_gelu").to(dev)
    layer.precision = 1
    layer.norm_neigh_dist_.fill_(1.0 / r)
    layer.norm_num_neighs_.fill_(sizes[t] / max(e, 1))
    x = torch.randn(sizes[s] * F, cin, device=dev, requires_grad=True)
    dy = torch.randn(sizes[t] * F, cout, device=dev)
    tf = tb = 0.0
    iters = 5
    for it in range(iters + 2):
        flush.fill_(it)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        y = layer(pin, pout, x, nb)
        ev[1].record()
        y.backward(dy)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            tf += ev[0].elapsed_time(ev[1]) / iters
            tb += ev[1].elapsed_time(ev[2]) / iters
        x.grad = None
        layer.zero_grad()
    tot += tf + tb
    out["layers"].append({"src": s, "dst": t, "radius": r, "edges": e, "cin": cin, "cout": cout, "fwd_ms": round(tf, 4),
                          "bwd_ms": round(tb, 4)})
out["conv_fwd_bwd_ms_total"] = round(tot, 3)
out["points_per_s_build_plus_convs"] = round(p.shape[0] / ((out["hierarchy_build_ms_median"] + tot) * 1e-3))
print(json.dumps(out))
