"""Checks se3_gemm_bf16_tn (tcgen05/TMEM and mma.sync implementations) against torch on a few shapes
and times them.  Run under `timeout`: a broken mbarrier protocol hangs instead of failing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from se3conv3d_b200 import _lib  # noqa: E402


def run(m, n, k, impl, out_bf16, alpha=0.5):
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(m * 31 + n * 7 + k)
    a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(n, k, device="cuda", generator=g).to(torch.bfloat16)
    c = torch.full((m, n), float("nan"), device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    _lib.check(L.se3_gemm_bf16_tn(_lib.ptr(a), _lib.ptr(b), m, n, k, alpha, _lib.ptr(c), int(out_bf16), impl,
                                  _lib.stream()), "se3_gemm_bf16_tn")
    torch.cuda.synchronize()
    ref = alpha * (a.double() @ b.double().t())
    err = float((c.double() - ref).abs().max() / ref.abs().max())
    return err, a, b, c


def main():
    shapes = [(128, 64, 64), (128, 64, 1024), (300, 32, 1024), (13780, 64, 1024), (1000, 16, 8), (257, 256, 2048),
              (4096, 1024, 64), (5000, 48, 520), (441000, 32, 1024)]
    ok = True
    for impl in (2, 1):
        for (m, n, k) in shapes:
            for ob in (False, True):
                err, a, b, c = run(m, n, k, impl, ob)
                tol = 1.5e-2 if ob else 2e-5
                flag = "ok" if err < tol else "FAIL"
                ok &= err < tol
                # timing
                L = _lib.lib()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for _ in range(2):
                    L.se3_gemm_bf16_tn(_lib.ptr(a), _lib.ptr(b), m, n, k, 0.5, _lib.ptr(c), int(ob), impl, _lib.stream())
                e0.record()
                for _ in range(10):
                    L.se3_gemm_bf16_tn(_lib.ptr(a), _lib.ptr(b), m, n, k, 0.5, _lib.ptr(c), int(ob), impl, _lib.stream())
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print("impl=%d m=%d n=%d k=%d out_bf16=%d  rel_err=%.2e %s  %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (
                    impl, m, n, k, ob, err, flag, ms, 2.0 * m * n * k / ms / 1e9,
                    (m * k * 2 + n * k * 2 + m * n * (2 if ob else 4)) / ms / 1e6), flush=True)
    print("GEMM_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
