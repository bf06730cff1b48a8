"""Single-layer micro benchmark / profiling target (BASELINE config 1 and the sweep shapes).

    python tools/layer_bench.py --n 8192 --radius 0.1 --frames 2 --cin 32 --cout 64 --precision 1 --iters 20
Prints per-iteration forward / backward times (CUDA events, L2 flushed between iterations)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200.layers import PNEConvLayerRotEquiv  # noqa: E402
from se3conv3d_b200.pc import PointcloudRotEquiv, BQNeighborhood  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--radius", type=float, default=0.1)
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--cin", type=int, default=32)
    ap.add_argument("--cout", type=int, default=64)
    ap.add_argument("--precision", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--density-scale", action="store_true", help="cube side ~ n^(1/3) so k-bar stays constant")
    a = ap.parse_args()
    dev = "cuda:0"
    side = (a.n / 8192.0) ** (1.0 / 3.0) if a.density_scale else 1.0
    pts = torch.rand(a.n, 3, generator=torch.Generator().manual_seed(0)) * side
    cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False,
           "n_frames": a.frames}
    pc = PointcloudRotEquiv(pts.to(dev), torch.zeros(a.n, dtype=torch.int32, device=dev), cfg)
    neigh = BQNeighborhood(pc, pc, a.radius)
    e = neigh.neighbors_.shape[0]
    torch.manual_seed(2)
    layer = PNEConvLayerRotEquiv(9, a.cin, a.cout, 32, "mlp_gelu").to(dev)
    layer.precision = a.precision
    layer.norm_neigh_dist_.fill_(1.0 / a.radius)
    layer.norm_num_neighs_.fill_(a.n / e)
    x = torch.randn(a.n * a.frames, a.cin, device=dev, requires_grad=True)
    dy = torch.randn(a.n * a.frames, a.cout, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tf = tb = 0.0
    for it in range(a.warmup + a.iters):
        flush.fill_(0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        y = layer(pc, pc, x, neigh)
        ev[1].record()
        y.backward(dy)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= a.warmup:
            tf += ev[0].elapsed_time(ev[1])
            tb += ev[1].elapsed_time(ev[2])
        x.grad = None
        layer.zero_grad()
    tf, tb = tf / a.iters, tb / a.iters
    # SURVEY 8(d) gather model, fp32 features: algorithmic bytes of one forward + backward
    kbar, f, w = e / a.n, a.frames, (a.cin * 32 * a.cout + 9 * 32 + 32) * 4
    fwd_b = a.n * (kbar * (16 + 36 * f + f * a.cin * 4) + (16 + 36 * f) + f * a.cout * 4) + w
    bwd_b = a.n * (kbar * (2 * (16 + 36 * f) + f * a.cin * 4 + f * a.cout * 4) + f * a.cout * 4 + f * a.cin * 4) + 2 * w
    print(json.dumps({"n": a.n, "edges": e, "kbar": round(e / a.n, 2), "frames": a.frames, "cin": a.cin, "cout": a.cout,
                      "precision": a.precision, "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4),
                      "points_per_s": round(a.n / ((tf + tb) * 1e-3)),
                      "alg_gbs_fwd": round(fwd_b / (tf * 1e-3) / 1e9, 1), "alg_gbs_fwd_bwd": round((fwd_b + bwd_b) / ((tf + tb) * 1e-3) / 1e9, 1)}))


if __name__ == "__main__":
    main()
