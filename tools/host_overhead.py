"""Host-side issue time of the dfaust conv stack (forward of the 21 layers + one backward sweep) against its device time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import workloads as wl, _lib  # noqa: E402

dev = torch.device("cuda:0")
pts, b = wl.synthetic_bodies(32, 6890, seed=0)
step = wl.DfaustStep(dev, precision=1)
pcs, neighs = step.build_hierarchy(pts.to(dev), b.to(dev), n_batches=32)
step.calibrate(pcs, neighs)
xs, dys = step.make_inputs(pcs)
for _ in range(5):
    step.conv_fwd_bwd(pcs, neighs)
    step.zero_grad()
torch.cuda.synchronize()
host_f, host_b, tot = [], [], []
for _ in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ys = [layer(pcs[li], pcs[lo], x, nb) for layer, nb, (_, li, lo, _, _, _), x in zip(step.layers, neighs, step.specs, xs)]
    t1 = time.perf_counter()
    torch.autograd.backward(ys, list(dys))
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    host_f.append(t1 - t0); host_b.append(t2 - t1); tot.append(t3 - t0)
    step.zero_grad()
med = lambda v: sorted(v)[len(v) // 2] * 1e3
print("host issue: forward x21 %.3f ms, backward sweep %.3f ms; wall incl. sync %.3f ms" % (med(host_f), med(host_b), med(tot)))
l0 = _lib.launch_count()
step.conv_fwd_bwd(pcs, neighs)
print("own launches per stack:", _lib.launch_count() - l0)
# one layer, python overhead per call
layer, nb, x = step.layers[7], neighs[7], xs[7]
li, lo = step.specs[7][1], step.specs[7][2]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    with torch.no_grad():
        layer(pcs[li], pcs[lo], x, nb)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("forward call host time (no grad): %.1f us" % ((t1 - t0) / 200 * 1e6))
