"""Times / checks the K-major projection GEMM implementations on the dfaust shapes: impl 1 = mma.sync, 2 = tcgen05 with
cp.async operand loads (round 1), 3 = persistent tcgen05 with TMA operand loads (round 2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200._lib import lib, check, ptr, stream  # noqa: E402

dev = "cuda:0"
L = lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(84468, 32, 1024, False), (44884, 32, 1024, False), (44884, 64, 2048, False), (84468, 1024, 32, True),
          (11522, 128, 4096, False), (2028, 256, 8192, False), (370, 256, 8192, False), (441000, 32, 1024, False),
          (44884, 2048, 64, True), (300, 48, 520, False), (5000, 16, 8, False)]
for (m, n, k, ob) in shapes:
    g = torch.Generator().manual_seed(m + n + k)
    a = (torch.randn(m, k, generator=g) / k ** 0.5).to(dev).to(torch.bfloat16).contiguous()
    b = torch.randn(n, k, generator=g).to(dev).to(torch.bfloat16).contiguous()
    ref = (a.double() @ b.double().t()) * 0.5
    row = "%7d x %5d x %5d %s" % (m, n, k, "bf16" if ob else "fp32")
    for impl in (2, 3):
        c = torch.empty((m, n), dtype=torch.bfloat16 if ob else torch.float32, device=dev)
        try:
            check(L.se3_gemm_bf16_tn(ptr(a), ptr(b), m, n, k, 0.5, ptr(c), 1 if ob else 0, impl, stream()), "gemm")
        except Exception as e:
            row += "  impl %d: %s" % (impl, str(e)[:40])
            continue
        torch.cuda.synchronize()
        err = float((c.double() - ref).abs().max() / ref.abs().max())
        ts = []
        for _ in range(7):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.se3_gemm_bf16_tn(ptr(a), ptr(b), m, n, k, 0.5, ptr(c), 1 if ob else 0, impl, stream())
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        gb = (2 * m * k + 2 * n * k + (2 if ob else 4) * m * n) / 1e9
        row += "  impl %d: %7.1f us %6.0f GB/s err %.1e" % (impl, t * 1e3, gb / (t * 1e-3), err)
    print(row, flush=True)
