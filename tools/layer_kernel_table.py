"""Per-layer device time of the three gather kernels + the whole fwd / bwd call (CUDA events), dfaust stack."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import _lib, workloads as wl  # noqa: E402

dev = torch.device("cuda:0")
pts, b = wl.synthetic_bodies(32, 6890, 0)
step = wl.DfaustStep(dev, precision=1)
pcs, neighs = step.build_hierarchy(pts.to(dev), b.to(dev), n_batches=32)
step.calibrate(pcs, neighs)
xs, dys = step.make_inputs(pcs)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("%-12s %7s %8s %4s %4s | %8s %8s %8s | %8s %8s" % ("layer", "M", "E", "cin", "cout", "agg_fwd", "agg_tr", "edge", "fwd_us", "bwd_us"))
tot = [0.0] * 5
for i, (name, li, lo, _, cin, cout) in enumerate(step.specs):
    layer, nb = step.layers[i], neighs[i]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0

    def one():
        global tf, tb
        flush.fill_(1)
        ev[0].record()
        y = layer(pcs[li], pcs[lo], xs[i], nb)
        ev[1].record()
        y.backward(dys[i])
        ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1])
        tb += ev[1].elapsed_time(ev[2])
        step.zero_grad()
    for _ in range(2):
        one()
    tf = tb = 0.0
    prof = _lib.profile_kernels(one, reps=5)
    g = nb.conv_geometry(pcs[li], pcs[lo])
    vals = [prof[0][1] * 1e3, prof[1][1] * 1e3, prof[2][1] * 1e3, tf / 5 * 1e3, tb / 5 * 1e3]
    tot = [a + v for a, v in zip(tot, vals)]
    print("%-12s %7d %8d %4d %4d | %8.1f %8.1f %8.1f | %8.1f %8.1f" % ((name, g.n_out, g.n_edges, cin, cout) + tuple(vals)))
print("%-12s %7s %8s %4s %4s | %8.1f %8.1f %8.1f | %8.1f %8.1f" % (("total", "", "", "", "") + tuple(tot)))
