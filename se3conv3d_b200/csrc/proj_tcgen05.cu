// tcgen05 / TMEM projection GEMM (sm_100a):  C[M,N] = alpha * A[M,K] * B[N,K]^T
//   A, B bf16, K-major (row-major [rows][K]);  C fp32 or bf16;  fp32 accumulation in tensor memory.
//
// Used for the three K-major x K-major products of the layer:
//   y  = s * T  . W        (A = T  [R, Cin*K],     B = W^T  [Cout, Cin*K])
//   dT = s * dy . W^T      (A = dy [R, Cout],      B = W    [Cin*K, Cout])
//   dx = s * U  . Wp^T     (A = U  [N*F, Cout*K],  B = Wp   [Cin, Cout*K])
//
// Structure (one CTA = one 128 x BN output tile, 128 threads):
//   * operands are streamed with 16-byte cp.async into a 4-stage ring of 128B-swizzled K-major tiles
//     (the canonical UMMA "SW128 K-major" layout: 128-byte rows, 16-byte chunk index XOR row%8,
//     8-row groups 1024 bytes apart), zero-filled outside the matrix;
//   * after a stage has landed (cp.async.wait_group + fence.proxy.async + barrier) ONE thread issues
//     four tcgen05.mma (kind::f16, M=128, N=BN, K=16) whose accumulator lives in TMEM, then
//     tcgen05.commit arrives on the stage's mbarrier so the ring slot can be refilled;
//   * the epilogue reads the accumulator with tcgen05.ld (32 lanes x 32 bit, 8 columns at a time),
//     scales it and stores rows straight to global memory.
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (commit), LDGSTS (cp.async).
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace se3 {

void splitk_reduce_launch(const float* partials, int splits, int64_t mn, float alpha, float* c, cudaStream_t st);

namespace {

constexpr int BM = 128;
constexpr int BK = 64;       // 64 bf16 = 128 bytes = one swizzle span
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * 128;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (ignored for swizzled K-major, set to 1),
//   [32,46) SBO >> 4 (1024 B between 8-row groups), [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive accumulator columns of this thread's TMEM lane (no wait: pair with tmem_ld_wait)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BN, bool OUT_BF16>
__global__ void __launch_bounds__(128) k_gemm_tcgen05(int M, int N, int K, float alpha, const __nv_bfloat16* __restrict__ A,
                                                      int64_t lda, const __nv_bfloat16* __restrict__ B, int64_t ldb,
                                                      void* __restrict__ Cout, int64_t ldc, int kb_per_split,
                                                      int64_t split_stride, int nstages) {
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bars[STAGES + 1];
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 1024-byte aligned operand ring
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  // nstages (<= STAGES) ring slots are allocated: short-K products keep their footprint small enough for
  // several CTAs per SM (their time is the epilogue, which needs the occupancy)
  const uint32_t a_ring = ring, b_ring = ring + nstages * A_STAGE_BYTES;
  // Row tiles are walked from the LAST to the first.  The A operand of the forward projection / dx product was just
  // written, front to back, by the aggregation kernel, and the tail of a tensor larger than L2 is what is still
  // resident: reading it back to front turns that tail into L2 hits instead of evicting it first.  The dT product
  // writes back to front for the same reason (the edge-gradient kernel reads dT front to back right after).
  const int bm = (int)(gridDim.y - 1 - blockIdx.y) * BM, bn = blockIdx.x * BN;
  // split-K: blockIdx.z owns k-blocks [kb0, kb0 + nkb) and writes its partial tile to Cout + z * split_stride
  const int nkb_all = (K + BK - 1) / BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int nkb = min(kb_per_split, nkb_all - kb0);
  if (split_stride) Cout = reinterpret_cast<float*>(Cout) + (int64_t)blockIdx.z * split_stride;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i <= STAGES; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // barriers and tensor memory are set up; everything below touches global memory
  const uint32_t tmem_d = tmem_base_slot;
  constexpr uint32_t IDESC = make_idesc(BM, BN);

  auto load_stage = [&](int stage, int kb) {
    const int k0 = kb * BK;
    const uint32_t as = a_ring + stage * A_STAGE_BYTES;
    const uint32_t bs = b_ring + stage * B_STAGE_BYTES;
#pragma unroll
    for (int i = 0; i < BM * 8 / 128; ++i) {
      const int id = tid + i * 128;
      const int r = id >> 3, c = id & 7;
      const bool p = (bm + r < M) && (k0 + c * 8 < K);
      const __nv_bfloat16* src = p ? A + (int64_t)(bm + r) * lda + k0 + c * 8 : A;
      cp_async16(as + r * 128 + ((c ^ (r & 7)) << 4), src, p);
    }
#pragma unroll
    for (int i = 0; i < BN * 8 / 128; ++i) {
      const int id = tid + i * 128;
      const int r = id >> 3, c = id & 7;
      const bool p = (bn + r < N) && (k0 + c * 8 < K);
      const __nv_bfloat16* src = p ? B + (int64_t)(bn + r) * ldb + k0 + c * 8 : B;
      cp_async16(bs + r * 128 + ((c ^ (r & 7)) << 4), src, p);
    }
  };

  for (int it = 0; it < nkb + nstages - 1; ++it) {
    if (it < nkb) {
      const int s = it % nstages;
      if (it >= nstages) mbar_wait(smem_u32(&bars[s]), (uint32_t)((it / nstages - 1) & 1));  // slot consumed by the MMA
      load_stage(s, kb0 + it);
    }
    cp_async_commit();
    const int c = it - (nstages - 1);
    if (c >= 0) {
      // k-block c has landed (for this thread's copies): at most nstages - 1 younger groups may be pending
      switch (nstages) {
        case 1: cp_async_wait<0>(); break;
        case 2: cp_async_wait<1>(); break;
        case 3: cp_async_wait<2>(); break;
        default: cp_async_wait<STAGES - 1>(); break;
      }
      fence_proxy_async();          // make them visible to the tensor-core (async) proxy
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const int s = c % nstages;
        const uint32_t as = a_ring + s * A_STAGE_BYTES;
        const uint32_t bs = b_ring + s * B_STAGE_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < BK / 16; ++k4) {
          // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle span
          umma_f16(tmem_d, make_desc_sw128(as + k4 * 32), make_desc_sw128(bs + k4 * 32), IDESC, (c > 0 || k4 > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars[s]));
        if (c == nkb - 1) umma_commit(smem_u32(&bars[STAGES]));
      }
    }
  }
  // accumulator complete
  mbar_wait(smem_u32(&bars[STAGES]), 0u);
  tc_fence_after();
  const int row = bm + warp * 32 + lane;
  if (OUT_BF16 && BN >= 64) {
    // bf16 output, wide tiles: 64 columns at a time through a warp-private, XOR-swizzled staging tile in the
    // (now idle) operand ring, so that every global store instruction writes four full 128-byte lines
    // instead of 32 scattered 16-byte pieces.
    unsigned char* stage = smem_dyn + (ring - raw) + warp * 4096;  // [32 rows][128 B]
    __nv_bfloat16* Cb = reinterpret_cast<__nv_bfloat16*>(Cout);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 64) {
      uint32_t r0[32], r1[32];
      tmem_ld32_nowait(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r0);
      tmem_ld32_nowait(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c0 + 32), r1);
      tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint32_t* src = ch < 4 ? r0 + ch * 8 : r1 + (ch - 4) * 8;
        uint4 p;
        p.x = pack_bf16(alpha * __uint_as_float(src[0]), alpha * __uint_as_float(src[1]));
        p.y = pack_bf16(alpha * __uint_as_float(src[2]), alpha * __uint_as_float(src[3]));
        p.z = pack_bf16(alpha * __uint_as_float(src[4]), alpha * __uint_as_float(src[5]));
        p.w = pack_bf16(alpha * __uint_as_float(src[6]), alpha * __uint_as_float(src[7]));
        *reinterpret_cast<uint4*>(stage + lane * 128 + ((ch ^ (lane & 7)) << 4)) = p;
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), ch = lane & 7;
        const int grow = bm + warp * 32 + r, gcol = bn + c0 + ch * 8;
        if (grow < M && gcol < N)
          *reinterpret_cast<uint4*>(Cb + (int64_t)grow * ldc + gcol) =
              *reinterpret_cast<const uint4*>(stage + r * 128 + ((ch ^ (r & 7)) << 4));
      }
      __syncwarp();
    }
  } else
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 8) {
    float v[8];
    tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    const int col = bn + c0;
    if (row < M && col < N) {
      if (OUT_BF16) {
        __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(Cout) + (int64_t)row * ldc + col;
        if (col + 8 <= N && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
          uint4 p;
          p.x = pack_bf16(alpha * v[0], alpha * v[1]);
          p.y = pack_bf16(alpha * v[2], alpha * v[3]);
          p.z = pack_bf16(alpha * v[4], alpha * v[5]);
          p.w = pack_bf16(alpha * v[6], alpha * v[7]);
          *reinterpret_cast<uint4*>(C) = p;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (col + i < N) C[i] = __float2bfloat16(alpha * v[i]);
        }
      } else {
        float* C = reinterpret_cast<float*>(Cout) + (int64_t)row * ldc + col;
        if (col + 8 <= N && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
          *reinterpret_cast<float4*>(C) = make_float4(alpha * v[0], alpha * v[1], alpha * v[2], alpha * v[3]);
          *reinterpret_cast<float4*>(C + 4) = make_float4(alpha * v[4], alpha * v[5], alpha * v[6], alpha * v[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (col + i < N) C[i] = alpha * v[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

template <int BN, bool OB>
int launch_cfg(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
               int64_t ldb, void* c, int64_t ldc, int splits, float* partials, cudaStream_t st) {
  auto kern = k_gemm_tcgen05<BN, OB>;
  const int nkb = (int)((k + BK - 1) / BK);
  if (OB || partials == nullptr || splits < 1) splits = 1;
  int per = (nkb + splits - 1) / splits;
  splits = (nkb + per - 1) / per;
  // narrow tiles (the forward projection and dx: long K, N <= 64) run a 3-deep ring, so that three CTAs share an SM:
  // measured 49 / 31 / 56 us against 55 / 37 / 70 us with four stages on the dfaust level-0/1 shapes
  const int max_stages = BN <= 64 ? 3 : STAGES;
  int nstages = per < max_stages ? per : max_stages;
  {
    static const int forced = getenv("SE3_GEMM_STAGES") ? atoi(getenv("SE3_GEMM_STAGES")) : 0;  // tuning aid
    if (forced >= 1 && forced < nstages) nstages = forced;
  }
  // the bf16 epilogue stages 4 KB per warp in the ring: keep at least 16 KB of it
  size_t smem = (size_t)nstages * (A_STAGE_BYTES + BN * 128) + 1024;
  if (smem < 16384 + 1024) smem = 16384 + 1024;
  SE3_SMEM_ONCE(kern, (size_t)STAGES * (A_STAGE_BYTES + BN * 128) + 1024);
  {
    // the whole L1/shared array as shared memory: residency of these CTAs is a shared-memory question
    static bool carve = false;
    if (!carve) {
      SE3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      carve = true;
    }
  }
  dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)splits);
  if (splits > 1) {
    SE3_CUDA(launch_pdl(kern, grid, dim3(128), smem, st, (int)m, (int)n, (int)k, 1.0f, a, lda, b, ldb, partials, n, per, m * n, nstages));
    SE3_LAUNCH_CHECK();
    splitk_reduce_launch(partials, splits, m * n, alpha, reinterpret_cast<float*>(c), st);
    SE3_LAUNCH_CHECK();
  } else {
    SE3_CUDA(launch_pdl(kern, grid, dim3(128), smem, st, (int)m, (int)n, (int)k, alpha, a, lda, b, ldb, c, ldc, nkb, (int64_t)0, nstages));
    SE3_LAUNCH_CHECK();
  }
  return SE3_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// MN-major variant: C[M,N] = alpha * A^T * B with A stored [K][M] (M contiguous) and B stored [K][N] (N contiguous)
// -- the weight gradient dW[(c,k), o] = sum_r T[r,(c,k)] dy[r,o], whose contraction index r is the ROW index of both
// operands.  Same ring / mbarrier / TMEM structure as k_gemm_tcgen05; the operand tiles are the canonical UMMA
// "MN-major SWIZZLE_128B" layout (cute::UMMA::make_umma_desc<Major::MN>): 128-byte lines of 64 consecutive M (or N)
// elements, one line per k, 8 lines = one 1024-byte swizzle atom (16-byte chunk index XOR line % 8), k-groups
// SBO = 1024 bytes apart, 64-element MN atoms LBO = 8 KB apart (all 8 k-groups of a stage).  A 16-byte cp.async
// moves 8 consecutive elements of one operand row -- exactly the global layout, no transposition anywhere.
// fp32 output only; split-K over the rows (blockIdx.z) with the ordered reduction of the K-major kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MN_BK = 64;                       // operand rows (k) per stage
constexpr uint32_t MN_LBO = (MN_BK / 8) * 1024; // bytes between 64-element MN atoms

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(MN_LBO >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor with both operands MN-major (a_major = bit 15, b_major = bit 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) { return make_idesc(m, n) | (1u << 15) | (1u << 16); }

template <int BN>
__global__ void __launch_bounds__(128) k_gemm_tcgen05_mn(int M, int N, int K, float alpha, const __nv_bfloat16* __restrict__ A,
                                                         int64_t lda, const __nv_bfloat16* __restrict__ B, int64_t ldb,
                                                         float* __restrict__ Cout, int64_t ldc, int kb_per_split,
                                                         int64_t split_stride, int nstages) {
  static_assert(BN % 64 == 0 && BN <= 256, "whole 64-element N atoms");
  constexpr int A_BYTES = BM * MN_BK * 2;  // 16 KB: two M atoms
  constexpr int B_BYTES = BN * MN_BK * 2;
  constexpr int TMEM_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bars[STAGES + 1];
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  const uint32_t a_ring = ring, b_ring = ring + nstages * A_BYTES;
  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  const int nkb_all = (K + MN_BK - 1) / MN_BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int nkb = min(kb_per_split, nkb_all - kb0);
  if (split_stride) Cout += (int64_t)blockIdx.z * split_stride;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i <= STAGES; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // barriers and tensor memory are set up; everything below touches global memory
  const uint32_t tmem_d = tmem_base_slot;
  constexpr uint32_t IDESC = make_idesc_mn(BM, BN);

  // element (k-row r, column col) of a tile -> atom col / 64, k-group r / 8, line r % 8, chunk (col % 64) / 8
  auto load_stage = [&](int stage, int kb) {
    const int k0 = kb * MN_BK;
    const uint32_t as = a_ring + stage * A_BYTES;
    const uint32_t bs = b_ring + stage * B_BYTES;
#pragma unroll
    for (int i = 0; i < MN_BK * (BM / 8) / 128; ++i) {
      const int id = tid + i * 128;
      const int r = id / (BM / 8), cc = id % (BM / 8), atom = cc >> 3, c = cc & 7;
      const bool p = (k0 + r < K) && (bm + cc * 8 < M);
      const __nv_bfloat16* src = p ? A + (int64_t)(k0 + r) * lda + bm + cc * 8 : A;
      cp_async16(as + atom * MN_LBO + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4), src, p);
    }
#pragma unroll
    for (int i = 0; i < MN_BK * (BN / 8) / 128; ++i) {
      const int id = tid + i * 128;
      const int r = id / (BN / 8), cc = id % (BN / 8), atom = cc >> 3, c = cc & 7;
      const bool p = (k0 + r < K) && (bn + cc * 8 < N);
      const __nv_bfloat16* src = p ? B + (int64_t)(k0 + r) * ldb + bn + cc * 8 : B;
      cp_async16(bs + atom * MN_LBO + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4), src, p);
    }
  };

  for (int it = 0; it < nkb + nstages - 1; ++it) {
    if (it < nkb) {
      const int s = it % nstages;
      if (it >= nstages) mbar_wait(smem_u32(&bars[s]), (uint32_t)((it / nstages - 1) & 1));
      load_stage(s, kb0 + it);
    }
    cp_async_commit();
    const int c = it - (nstages - 1);
    if (c >= 0) {
      switch (nstages) {
        case 1: cp_async_wait<0>(); break;
        case 2: cp_async_wait<1>(); break;
        case 3: cp_async_wait<2>(); break;
        default: cp_async_wait<STAGES - 1>(); break;
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const int s = c % nstages;
        const uint32_t as = a_ring + s * A_BYTES;
        const uint32_t bs = b_ring + s * B_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < MN_BK / 16; ++k4) {
          // advancing K by 16 = two 8-line k-groups = 2048 bytes
          umma_f16(tmem_d, make_desc_mn_sw128(as + k4 * 2048), make_desc_mn_sw128(bs + k4 * 2048), IDESC,
                   (c > 0 || k4 > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars[s]));
        if (c == nkb - 1) umma_commit(smem_u32(&bars[STAGES]));
      }
    }
  }
  mbar_wait(smem_u32(&bars[STAGES]), 0u);
  tc_fence_after();
  const int row = bm + warp * 32 + lane;
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 8) {
    float v[8];
    tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    const int col = bn + c0;
    if (row < M && col < N) {
      float* C = Cout + (int64_t)row * ldc + col;
      if (col + 8 <= N && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
        *reinterpret_cast<float4*>(C) = make_float4(alpha * v[0], alpha * v[1], alpha * v[2], alpha * v[3]);
        *reinterpret_cast<float4*>(C + 4) = make_float4(alpha * v[4], alpha * v[5], alpha * v[6], alpha * v[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (col + i < N) C[i] = alpha * v[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

template <int BN>
int launch_mn_cfg(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                  int64_t ldb, float* c, int64_t ldc, int splits, float* partials, cudaStream_t st) {
  auto kern = k_gemm_tcgen05_mn<BN>;
  const int nkb = (int)((k + MN_BK - 1) / MN_BK);
  if (partials == nullptr || splits < 1 || ldc != n) splits = 1;
  int per = (nkb + splits - 1) / splits;
  splits = (nkb + per - 1) / per;
  const int max_stages = BN <= 64 ? 4 : (BN <= 128 ? 4 : 3);
  const int nstages = per < max_stages ? per : max_stages;
  const size_t stage = (size_t)BM * MN_BK * 2 + (size_t)BN * MN_BK * 2;
  const size_t smem = nstages * stage + 1024;
  SE3_SMEM_ONCE(kern, (size_t)STAGES * stage + 1024);
  dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)splits);
  if (splits > 1) {
    SE3_CUDA(launch_pdl(kern, grid, dim3(128), smem, st, (int)m, (int)n, (int)k, 1.0f, a, lda, b, ldb, partials, n, per, m * n, nstages));
    SE3_LAUNCH_CHECK();
    splitk_reduce_launch(partials, splits, m * n, alpha, c, st);
    SE3_LAUNCH_CHECK();
  } else {
    SE3_CUDA(launch_pdl(kern, grid, dim3(128), smem, st, (int)m, (int)n, (int)k, alpha, a, lda, b, ldb, c, ldc, nkb, (int64_t)0, nstages));
    SE3_LAUNCH_CHECK();
  }
  return SE3_OK;
}

}  // namespace

bool tcgen05_gemm_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb) {
  return m >= 1 && n >= 16 && (n % 16) == 0 && (k % 8) == 0 && (lda % 8) == 0 && (ldb % 8) == 0;
}

// Split-K factor for a (fp32-output, ldc == n) product whose output tiles cannot fill the GPU: the k-blocks are
// spread over enough CTAs for ~2 per SM, at least 4 k-blocks each.
int tcgen05_gemm_splits(int64_t m, int64_t n, int64_t k) {
  const int bn = n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)));
  const int64_t tiles = ((m + BM - 1) / BM) * ((n + bn - 1) / bn);
  const int64_t nkb = (k + BK - 1) / BK;
  if (tiles >= num_sms() || nkb < 8) return 1;
  int64_t s = (2 * (int64_t)num_sms() + tiles - 1) / tiles;
  if (s > nkb / 4) s = nkb / 4;
  if (s > 64) s = 64;
  return s < 1 ? 1 : (int)s;
}

// C = alpha * A[M,K] . B[N,K]^T ; out_bf16 selects the output type.  splits > 1 (fp32 output with ldc == n only)
// runs split-K through `partials` (splits * m * n floats) and an ordered reduction.
int launch_gemm_tcgen05(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda,
                        const __nv_bfloat16* b, int64_t ldb, void* c, int64_t ldc, bool out_bf16, int splits,
                        float* partials, cudaStream_t st) {
  if (ldc != n) splits = 1;
  if (!tcgen05_gemm_supported(m, n, k, lda, ldb)) {
    set_error("launch_gemm_tcgen05: unsupported shape m=%lld n=%lld k=%lld", (long long)m, (long long)n, (long long)k);
    return SE3_EINVAL;
  }
  // one N tile when N <= 256, else 256/128-wide tiles (N % 16 == 0 is guaranteed)
  const int bn = n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)));
#define SE3_TC_CASE(BN_)                                                                                           \
  case BN_:                                                                                                         \
    return out_bf16 ? launch_cfg<BN_, true>(m, n, k, alpha, a, lda, b, ldb, c, ldc, 1, nullptr, st)                 \
                    : launch_cfg<BN_, false>(m, n, k, alpha, a, lda, b, ldb, c, ldc, splits, partials, st);
  switch (bn) {
    SE3_TC_CASE(16)
    SE3_TC_CASE(32)
    SE3_TC_CASE(64)
    SE3_TC_CASE(128)
    SE3_TC_CASE(256)
  }
#undef SE3_TC_CASE
  return SE3_EINVAL;
}

// C[M,N] = alpha * A^T B, A stored [K][M], B stored [K][N] (bf16, fp32 output): the weight gradient on tcgen05.
bool tcgen05_gemm_mn_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb) {
  // N need not fill its last 64-column atom: the missing columns are zero-filled in shared memory and never stored
  return m >= 1 && k >= 1 && n >= 8 && (n % 8) == 0 && (m % 8) == 0 && (lda % 8) == 0 && (ldb % 8) == 0;
}
int launch_gemm_tcgen05_mn(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda,
                           const __nv_bfloat16* b, int64_t ldb, float* c, int64_t ldc, int splits, float* partials,
                           cudaStream_t st) {
  if (!tcgen05_gemm_mn_supported(m, n, k, lda, ldb)) {
    set_error("launch_gemm_tcgen05_mn: unsupported shape m=%lld n=%lld k=%lld", (long long)m, (long long)n, (long long)k);
    return SE3_EINVAL;
  }
  const int bn = n <= 64 ? 64 : (n <= 128 ? 128 : 256);
  switch (bn) {
    case 64: return launch_mn_cfg<64>(m, n, k, alpha, a, lda, b, ldb, c, ldc, splits, partials, st);
    case 128: return launch_mn_cfg<128>(m, n, k, alpha, a, lda, b, ldb, c, ldc, splits, partials, st);
    default: return launch_mn_cfg<256>(m, n, k, alpha, a, lda, b, ldb, c, ldc, splits, partials, st);
  }
}

}  // namespace se3
