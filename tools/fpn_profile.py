"""Host- and device-side profile of the full FPN training step (torch.profiler): op counts, top ops by host time,
kernel time by name.  usage: python tools/fpn_profile.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from se3conv3d_b200 import workloads as wl  # noqa: E402
import stage_reference_models  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
seg_models = stage_reference_models.import_models("seg_models")
pts, b = wl.synthetic_bodies(32, 6890, seed=0)
pts_d, b_d = pts.to(dev), b.to(dev)
labels = torch.randint(0, 20, (pts.shape[0],), device=dev)
fs = wl.FpnStep(dev, seg_models, precision=1)
for _ in range(4):
    fs.step(pts_d, b_d, labels, 32)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        fs.step(pts_d, b_d, labels, 32)
    torch.cuda.synchronize()
ka = prof.key_averages()
print("==== top ops by self CPU time (per step)")
rows = sorted(ka, key=lambda e: -e.self_cpu_time_total)[:28]
for e in rows:
    print("%-60s n=%6.1f  self cpu %8.1f us  cuda %8.1f us" % (e.key[:60], e.count / steps, e.self_cpu_time_total / steps,
                                                               getattr(e, "self_device_time_total", 0) / steps))
print("==== top kernels by device time (per step)")
rows = sorted(ka, key=lambda e: -getattr(e, "self_device_time_total", 0))[:28]
for e in rows:
    print("%-70s n=%6.1f  %8.1f us" % (e.key[:70], e.count / steps, getattr(e, "self_device_time_total", 0) / steps))
tot_k = sum(getattr(e, "self_device_time_total", 0) for e in ka) / steps
print("device time per step %.2f ms" % (tot_k / 1e3))
