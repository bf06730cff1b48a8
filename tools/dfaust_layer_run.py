"""Runs ONE layer of the dfaust conv stack (by name) forward+backward a few times on the real hierarchy;
the target of `ncu` captures of the dominant kernels.

    python tools/dfaust_layer_run.py seg_head 3 1      # name, iterations, precision
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import workloads as wl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "seg_head"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
precision = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
pts, b = wl.synthetic_bodies(32, 6890, 0)
step = wl.DfaustStep(dev, precision=precision)
pcs, neighs = step.build_hierarchy(pts.to(dev), b.to(dev))
step.calibrate(pcs, neighs)
xs, dys = step.make_inputs(pcs)
idx = [s[0] for s in step.specs].index(name)
layer, nb, (_, li, lo, _, cin, cout) = step.layers[idx], neighs[idx], step.specs[idx]
x, dy = xs[idx], dys[idx]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(iters):
    flush.fill_(it)
    ev[0].record()
    y = layer(pcs[li], pcs[lo], x, nb)
    ev[1].record()
    y.backward(dy)
    ev[2].record()
    torch.cuda.synchronize()
    step.zero_grad()
print("%s: M=%d E=%d %d->%d  fwd %.3f ms  bwd %.3f ms" % (name, pcs[lo].pts_.shape[0], nb.conv_geometry(pcs[li], pcs[lo]).n_edges, cin, cout,
                                                      ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
