"""Runs the dfaust conv stack (fwd+bwd x21) a few times; target for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import workloads as wl  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
precision = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
pts, b = wl.synthetic_bodies(32, 6890, 0)
step = wl.DfaustStep(dev, precision=precision)
pcs, neighs = step.build_hierarchy(pts.to(dev), b.to(dev), n_batches=32)
step.calibrate(pcs, neighs)
step.make_inputs(pcs)
torch.cuda.synchronize()
for _ in range(iters):
    step.conv_fwd_bwd(pcs, neighs)
    step.zero_grad()
torch.cuda.synchronize()
print("done")
