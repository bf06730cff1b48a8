"""Host-side profile of one hierarchy build (cProfile, steady state) + wall/GPU split."""
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
pts, b = pts.to(dev), b.to(dev)
step = wl.DfaustStep(dev, precision=1)
for _ in range(10):
    pcs, neighs = step.build_hierarchy(pts, b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    pcs, neighs = step.build_hierarchy(pts, b)
th = time.perf_counter() - t0
torch.cuda.synchronize()
print("hierarchy build: host %.2f ms, total %.2f ms" % (th / 10 * 1e3, (time.perf_counter() - t0) / 10 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    pcs, neighs = step.build_hierarchy(pts, b)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(30)
st.sort_stats("cumulative").print_stats(40)
