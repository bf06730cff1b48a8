"""Times the tcgen05 projection GEMM on the dfaust seg_head shapes with the L2 flushed before every launch
(SE3_GEMM_STAGES=n caps the operand ring depth: tuning aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import _lib  # noqa: E402

L = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (m, n, k, ob) in [(84468, 32, 1024, 0), (84468, 1024, 32, 1), (44884, 32, 1024, 0), (44884, 64, 2048, 0), (11522, 128, 4096, 0)]:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    c = torch.empty((m, n), device="cuda", dtype=torch.bfloat16 if ob else torch.float32)
    ts = []
    for it in range(8):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.se3_gemm_bf16_tn(_lib.ptr(a), _lib.ptr(b), m, n, k, 0.5, _lib.ptr(c), ob, 2, _lib.stream())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    ms = ts[len(ts) // 2]
    print("stages=%s m=%d n=%d k=%d bf16out=%d  %.1f us  %.0f GB/s" % (os.environ.get("SE3_GEMM_STAGES", "auto"), m, n, k, ob, ms * 1e3,
          (m * k * 2 + n * k * 2 + m * n * (2 if ob else 4)) / ms / 1e6), flush=True)
