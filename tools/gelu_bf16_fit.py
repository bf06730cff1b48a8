"""Error study behind the packed bf16x2 GELU / GELU' of the tensor-core path (csrc/tc_common.cuh).

Emulates bf16 round-to-nearest-even arithmetic (mul.rn / fma.rn.bf16x2, one rounding per instruction) in numpy and
compares, against the exact erf GELU in fp64:
  * the fp32 tanh-form evaluation followed by one bf16 rounding (the round-1 kernels), and
  * the evaluation carried out in bf16 throughout, for the constants on the bf16 grid around (2 c0, 8 c1).
Prints rms / max / mean error for pre-activations ~ N(0, 1.5^2).  Run on the build box: no GPU needed.
"""
import math

import numpy as np
from scipy.special import erf

C0, C1 = 0.8000095, 0.03476866


def bf(x):
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def fma_bf(a, b, c):
    return bf(a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64))


def mul_bf(a, b):
    return bf(a.astype(np.float64) * b.astype(np.float64))


def packed(y, a_c, b_c, da_c, db_c):
    full = lambda v: np.full_like(y, v)  # noqa: E731
    yb = bf(y)
    y2 = mul_bf(yb, yb)
    p = fma_bf(full(b_c), y2, full(a_c))
    th = bf(np.tanh(mul_bf(yb, p).astype(np.float64)))
    out = fma_bf(yb, th, yb)
    s = fma_bf(-th, th, full(1.0))
    cdf = fma_bf(full(0.5), th, full(0.5))
    dp = fma_bf(full(db_c), y2, full(da_c))
    return out, fma_bf(mul_bf(yb, s), dp, cdf)


def stats(name, v, r):
    e = v.astype(np.float64) - r
    print("%-34s rms %.3e  max %.3e  rel-rms %.3e  mean %.3e" % (
        name, np.sqrt((e ** 2).mean()), np.abs(e).max(), np.sqrt((e ** 2).mean()) / np.sqrt((r ** 2).mean()), e.mean()))


def main():
    rng = np.random.default_rng(0)
    x = rng.normal(0, 1.5, 2_000_000).astype(np.float32)
    y = (0.5 * x).astype(np.float32)
    xd = x.astype(np.float64)
    ref = 0.5 * xd * (1 + erf(xd / math.sqrt(2)))
    refg = 0.5 * (1 + erf(xd / math.sqrt(2))) + xd * np.exp(-xd ** 2 / 2) / math.sqrt(2 * math.pi)
    y2 = y * y
    th = np.tanh(y * (8 * C1 * y2 + 2 * C0))
    stats("fp32 tanh form -> bf16: gelu", bf(y + y * th), ref)
    stats("fp32 tanh form -> bf16: gelu'", bf((y * (1 - th * th)) * (12 * C1 * y2 + C0) + (.5 + .5 * th)), refg)
    o, g = packed(y, 1.6015625, 0.26953125, 0.80078125, 0.404296875)
    stats("bf16 throughout (shipped): gelu", o, ref)
    stats("bf16 throughout (shipped): gelu'", g, refg)
    a0 = int(bf(np.float32(2 * C0)).view(np.uint32))
    b0 = int(bf(np.float32(8 * C1)).view(np.uint32))
    for da in (-1, 0, 1):
        for db in range(-4, 5):
            a_c = np.array([a0 + (da << 16)], dtype=np.uint32).view(np.float32)[0]
            b_c = np.array([b0 + (db << 16)], dtype=np.uint32).view(np.float32)[0]
            o, _ = packed(y, a_c, b_c, a_c / 2, 1.5 * b_c)
            e = o.astype(np.float64) - ref
            print("A %.7f B %.8f  rms %.3e mean %+.3e" % (a_c, b_c, np.sqrt((e ** 2).mean()), e.mean()))


if __name__ == "__main__":
    main()
