"""GPU timeline of the pipelined hot-path step (torch.profiler / CUPTI): per-stream busy time, idle gaps of the device,
and the largest gaps with the kernels around them.  usage: python tools/timeline.py [steps] [--thread]"""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import workloads as wl  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4
threaded = "--thread" in sys.argv
dev = torch.device("cuda:0")
pts, b = wl.synthetic_bodies(32, 6890, seed=0)
pts_d, b_d = pts.to(dev), b.to(dev)
step = wl.DfaustStep(dev, precision=1)
pcs, neighs = step.build_hierarchy(pts_d, b_d, n_batches=32)
step.calibrate(pcs, neighs)
step.make_inputs(pcs)
side = torch.cuda.Stream(dev, priority=-1)
step.run_pipelined([(pts_d, b_d)] * 6, 32, side, threaded=threaded)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step.run_pipelined([(pts_d, b_d)] * steps, 32, side, threaded=threaded)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
k = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
k.sort(key=lambda e: e["ts"])
t0, t1 = k[0]["ts"], max(e["ts"] + e["dur"] for e in k)
print("kernels %d, span %.3f ms (%.3f ms per step)" % (len(k), (t1 - t0) / 1e3, (t1 - t0) / 1e3 / steps))
by_stream = {}
for e in k:
    by_stream.setdefault(e["args"].get("stream"), []).append(e)
for s, lst in sorted(by_stream.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print("stream %s: %d kernels, busy %.3f ms" % (s, len(lst), sum(e["dur"] for e in lst) / 1e3))
# union of busy intervals over all streams -> idle gaps
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in k)
gaps, cur_end, busy = [], iv[0][1], 0.0
cur_start = iv[0][0]
for a, bnd in iv[1:]:
    if a > cur_end:
        gaps.append((a - cur_end, cur_end, a))
        busy += cur_end - cur_start
        cur_start = a
    cur_end = max(cur_end, bnd)
busy += cur_end - cur_start
print("device busy (any stream) %.3f ms, idle %.3f ms in %d gaps" % (busy / 1e3, sum(g[0] for g in gaps) / 1e3, len(gaps)))
hist = [0, 0, 0, 0]
for g in gaps:
    hist[0 if g[0] < 5 else 1 if g[0] < 20 else 2 if g[0] < 100 else 3] += g[0]
print("idle by gap size: <5us %.3f ms, 5-20us %.3f ms, 20-100us %.3f ms, >100us %.3f ms" % tuple(h / 1e3 for h in hist))
cpu = [e for e in ev if e.get("cat") == "cuda_runtime" and "dur" in e]
print("cuda runtime calls %d, host time in them %.3f ms" % (len(cpu), sum(e["dur"] for e in cpu) / 1e3))
names = {}
for e in cpu:
    n = names.setdefault(e["name"], [0, 0.0])
    n[0] += 1
    n[1] += e["dur"]
for n, (c, d) in sorted(names.items(), key=lambda kv: -kv[1][1])[:8]:
    print("   %-40s n=%5d  %.3f ms" % (n, c, d / 1e3))
for g in sorted(gaps, reverse=True)[:12]:
    before = max((e for e in k if e["ts"] + e["dur"] <= g[1] + 0.5), key=lambda e: e["ts"] + e["dur"])
    after = min((e for e in k if e["ts"] >= g[2] - 0.5), key=lambda e: e["ts"])
    print("gap %.1f us at %.3f ms: after %s | before %s" % (g[0], (g[1] - t0) / 1e3, before["name"][:50], after["name"][:50]))
