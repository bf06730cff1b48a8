"""Per-step device times of the pipelined hot-path loop (events after every conv stack) and allocator statistics --
looks for outlier steps.  usage: python tools/step_times.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import workloads as wl  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda:0")
pts, b = wl.synthetic_bodies(32, 6890, seed=0)
pts_d, b_d = pts.to(dev), b.to(dev)
step = wl.DfaustStep(dev, precision=1)
pcs, neighs = step.build_hierarchy(pts_d, b_d, n_batches=32)
step.calibrate(pcs, neighs)
step.make_inputs(pcs)
side = torch.cuda.Stream(dev, priority=-1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
step.run_pipelined([(pts_d, b_d)] * 10, 32, side)
torch.cuda.synchronize()
for rep in range(3):
    evs = []
    s0 = torch.cuda.memory_stats()

    def after(out, i):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        evs.append(e)
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    step.run_pipelined([(pts_d, b_d)] * steps, 32, side, before_conv=lambda: flush.fill_(1), after_conv=after)
    torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    ts = [e0.elapsed_time(evs[0])] + [evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]
    srt = sorted(ts)
    print("rep %d: mean %.3f ms  median %.3f  p90 %.3f  max %.3f (step %d)  min %.3f; cudaMalloc calls %d, frees %d, retries %d" % (
        rep, sum(ts) / len(ts), srt[len(ts) // 2], srt[int(0.9 * len(ts))], srt[-1], ts.index(srt[-1]), srt[0],
        s1["num_device_alloc"] - s0["num_device_alloc"], s1["num_device_free"] - s0["num_device_free"],
        s1["num_alloc_retries"] - s0["num_alloc_retries"]))
    print("   worst five:", ["%.2f" % v for v in srt[-5:]])
