"""One mid-size layer (40000 points, 32->32, F=2) forward + backward a few times: target of ncu captures and of the
per-kernel device timers."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from fused_check import problem  # noqa: E402
from se3conv3d_b200 import _lib  # noqa: E402

n = int(os.environ.get("N", "40000"))
cin = int(os.environ.get("CIN", "32"))
cout = int(os.environ.get("COUT", "32"))
pc, neigh, layer, x, dy = problem(n, 0.055 * (40000 / n) ** (1 / 3), 2, cin, cout)
xx = x.clone().requires_grad_(True)


def once():
    y = layer(pc, pc, xx, neigh)
    y.backward(dy)


for _ in range(2):
    once()
torch.cuda.synchronize()
reps = int(os.environ.get("REPS", "10"))
t0 = time.perf_counter()
for _ in range(reps):
    once()
torch.cuda.synchronize()
print("n=%d E=%d %d->%d fwd+bwd %.3f ms" % (n, neigh.neighbors_.shape[0], cin, cout, (time.perf_counter() - t0) * 1e3 / reps))
if not os.environ.get("NOPROF"):
    for name, ms, cnt in _lib.profile_kernels(once, reps=5):
        print("  %-50s %.1f us x %d" % (name, ms * 1e3, cnt))
