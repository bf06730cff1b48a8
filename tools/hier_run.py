"""Builds the fused dfaust hierarchy a few times (target of an ncu launch list)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import workloads as wl  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
pts, b = wl.synthetic_bodies(32, 6890, 0)
pts, b = pts.to(dev), b.to(dev)
step = wl.DfaustStep(dev, precision=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(iters):
    flush.fill_(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step.build_hierarchy(pts, b, n_batches=32)
    torch.cuda.synchronize()
    print("build %.3f ms" % ((time.perf_counter() - t0) * 1e3))
