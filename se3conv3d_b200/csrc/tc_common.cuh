// Tensor-core helpers shared by the precision-1 kernels: bf16 packing, ldmatrix, mma.sync m16n8k16,
// cp.async, and the cheap erf-GELU used on the tensor-core path.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace se3 {

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  // cvt.rn.bf16x2.f32 d, a, b : a -> upper half, b -> lower half
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

// D[16x8] += A[16x16] * B[16x8], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// erf-GELU through Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): two MUFU ops (rcp, ex2) and 11
// FMA-pipe ops, branch free, instead of erff's two-branch polynomial.
//   q = 0.5 erfc(|x|/sqrt2) = 0.5 t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-x^2/2),  t = 1/(1 + p |x|/sqrt2)
//   gelu(x) = relu(x) - |x| q          gelu'(x) = Phi(x) + x phi(x),  Phi = x >= 0 ? 1 - q : q
__device__ __forceinline__ float gelu_q(float x, float& e) {
  const float t = rcp_approx(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752440f, 1.0f));
  e = ex2_approx(x * x * -0.72134752044448170368f);  // exp(-x^2/2)
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  return p * t * e;
}

template <int ACT>
__device__ __forceinline__ float act_fast(float x) {
  if (ACT == 2) {
    float e;
    const float q = gelu_q(x, e);
    return fmaf(-fabsf(x), q, fmaxf(x, 0.0f));
  }
  if (ACT == 1) return fmaxf(x, 0.0f);
  if (ACT == 3) return __sinf(x);
  return x;
}
template <int ACT>
__device__ __forceinline__ float act_grad_fast(float x) {
  if (ACT == 2) {
    float e;
    const float q = gelu_q(x, e);
    const float cdf = x >= 0.0f ? 1.0f - q : q;
    return fmaf(x * 0.39894228040143267794f, e, cdf);
  }
  if (ACT == 1) return x > 0.0f ? 1.0f : 0.0f;
  if (ACT == 3) return __cosf(x);
  return 1.0f;
}

}  // namespace se3
