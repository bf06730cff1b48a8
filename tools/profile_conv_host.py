"""Host-side profile of the conv stack (21 fwd + one backward sweep) in steady state."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import workloads as wl  # noqa: E402

dev = torch.device("cuda:0")
pts, b = wl.synthetic_bodies(32, 6890, 0)
step = wl.DfaustStep(dev, precision=1)
pcs, neighs = step.build_hierarchy(pts.to(dev), b.to(dev), n_batches=32)
step.calibrate(pcs, neighs)
step.make_inputs(pcs)
for _ in range(5):
    step.conv_fwd_bwd(pcs, neighs)
    step.zero_grad()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step.conv_fwd_bwd(pcs, neighs)
    step.zero_grad()
th = time.perf_counter() - t0
torch.cuda.synchronize()
print("conv stack: host %.2f ms, total %.2f ms" % (th / 10 * 1e3, (time.perf_counter() - t0) / 10 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step.conv_fwd_bwd(pcs, neighs)
    step.zero_grad()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
