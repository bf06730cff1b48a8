// Tensor-core helpers shared by the precision-1 kernels: bf16 packing, ldmatrix, mma.sync m16n8k16,
// cp.async, and the single-MUFU erf-GELU used on the tensor-core path.
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

// D = A * B (no accumulator input)
__device__ __forceinline__ void mma_bf16_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}

// D[16x8] += A[16x8] * B[8x8], tf32 inputs, fp32 accumulate.
//   A: a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4);  B: b0=(k=t,n=g) b1=(k=t+4,n=g);  D as m16n8k16.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D = A * B (no accumulator input: the C operand is the zero register, nothing to clear beforehand)
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

// erf-GELU for the tensor-core path through ONE MUFU op and four FMA-pipe ops.
//   Phi(x) ~= 0.5 (1 + tanh(x (c0 + c1 x^2))), (c0, c1) = minimax fit to the erf form: max |gelu error| 2.9e-4,
//   max |gelu' error| 8.3e-4 -- at or below the bf16 rounding of the basis values they produce.
// The kernels feed y = x / 2 (the 1/2 is folded into the tensor-core operand of the basis affine), so
//   gelu(x) = y + y tanh(y (2 c0 + 8 c1 y^2)).
constexpr float GELU_C0 = 0.8000095f, GELU_C1 = 0.03476866f;
constexpr float GELU_PRE_SCALE = 0.5f;  // y = GELU_PRE_SCALE * x
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gelu_half_arg(float y) {
  const float y2 = y * y;
  const float p = fmaf(8.0f * GELU_C1, y2, 2.0f * GELU_C0);
  const float th = tanh_approx(y * p);
  return fmaf(y, th, y);
}
__device__ __forceinline__ float gelu_grad_half_arg(float y) {
  const float y2 = y * y;
  const float p = fmaf(8.0f * GELU_C1, y2, 2.0f * GELU_C0);
  const float dp = fmaf(12.0f * GELU_C1, y2, GELU_C0);  // (c0 + 3 c1 x^2), x = 2y
  const float th = tanh_approx(y * p);
  const float s = fmaf(-th, th, 1.0f);
  const float cdf = fmaf(0.5f, th, 0.5f);
  return fmaf(y * s, dp, cdf);
}

// ---- packed bf16x2 form of the same GELU (SE3_PACKED_ACT, default on): two values per FMA-pipe instruction
// (HFMA2.BF16_V2 / HMUL2.BF16_V2) and MUFU.TANH.BF16 -- 4 instructions per value instead of 5.5 (gelu) and 7 instead
// of 10.5 (gelu' times the incoming gradient).  The basis value is rounded to bf16 anyway (it is a tensor-core
// operand); evaluating in bf16 raises its rms error from 1.7e-3 to 2.6e-3 relative (gelu' 2.1e-3 -> 3.9e-3),
// constants re-fitted on the bf16 grid so that the mean error stays ~1e-5 (tools/gelu_bf16_fit.py).
#ifndef SE3_PACKED_ACT
#define SE3_PACKED_ACT 1
#endif
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t bf2_tanh(uint32_t a) {
  uint32_t r;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}
constexpr uint32_t BF2_GELU_A = 0x3FCD3FCDu;   // 1.6015625   ~ 2 c0
constexpr uint32_t BF2_GELU_B = 0x3E8A3E8Au;   // 0.26953125  ~ 8 c1 (re-fitted for A on the bf16 grid)
constexpr uint32_t BF2_GELU_DA = 0x3F4D3F4Du;  // 0.80078125  = A / 2
constexpr uint32_t BF2_GELU_DB = 0x3ECF3ECFu;  // 0.404296875 = 1.5 B
constexpr uint32_t BF2_HALF = 0x3F003F00u, BF2_ONE = 0x3F803F80u, BF2_NEG = 0x80008000u;
// y = (pre / 2) as a bf16 pair -> gelu(pre) as a bf16 pair
__device__ __forceinline__ uint32_t gelu_half_arg_bf2(uint32_t y) {
  const uint32_t p = bf2_fma(BF2_GELU_B, bf2_mul(y, y), BF2_GELU_A);
  return bf2_fma(y, bf2_tanh(bf2_mul(y, p)), y);
}
__device__ __forceinline__ uint32_t gelu_grad_half_arg_bf2(uint32_t y) {
  const uint32_t y2 = bf2_mul(y, y);
  const uint32_t p = bf2_fma(BF2_GELU_B, y2, BF2_GELU_A);
  const uint32_t dp = bf2_fma(BF2_GELU_DB, y2, BF2_GELU_DA);
  const uint32_t th = bf2_tanh(bf2_mul(y, p));
  const uint32_t s = bf2_fma(th ^ BF2_NEG, th, BF2_ONE);
  const uint32_t cdf = bf2_fma(BF2_HALF, th, BF2_HALF);
  return bf2_fma(bf2_mul(y, s), dp, cdf);
}

// gelu(pre) and gelu'(pre) from ONE tanh (the merged backward pass needs both for the same basis entry)
__device__ __forceinline__ void gelu_both_half_arg_bf2(uint32_t y, uint32_t& h, uint32_t& g) {
  const uint32_t y2 = bf2_mul(y, y);
  const uint32_t p = bf2_fma(BF2_GELU_B, y2, BF2_GELU_A);
  const uint32_t dp = bf2_fma(BF2_GELU_DB, y2, BF2_GELU_DA);
  const uint32_t th = bf2_tanh(bf2_mul(y, p));
  const uint32_t s = bf2_fma(th ^ BF2_NEG, th, BF2_ONE);
  const uint32_t cdf = bf2_fma(BF2_HALF, th, BF2_HALF);
  g = bf2_fma(bf2_mul(y, s), dp, cdf);
  h = bf2_fma(y, th, y);
}

// Activations of the tensor-core path take z = act_pre_scale(ACT) * pre (z = pre / 2 for GELU, pre otherwise).
__host__ __device__ __forceinline__ float act_pre_scale(int act) { return act == 2 ? GELU_PRE_SCALE : 1.0f; }

template <int ACT>
__device__ __forceinline__ float act_fast(float x) {
  if (ACT == 2) return gelu_half_arg(x);
  if (ACT == 1) return fmaxf(x, 0.0f);
  if (ACT == 3) return __sinf(x);
  return x;
}
template <int ACT>
__device__ __forceinline__ float act_grad_fast(float x) {
  if (ACT == 2) return gelu_grad_half_arg(x);
  if (ACT == 1) return x > 0.0f ? 1.0f : 0.0f;
  if (ACT == 3) return __cosf(x);
  return 1.0f;
}

}  // namespace se3
