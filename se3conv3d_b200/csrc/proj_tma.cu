// Persistent, warp-specialised tcgen05 GEMM fed by the TMA engine (sm_100a):  C[M,N] = alpha * A[M,K] * B[N,K]^T
//   A, B bf16 K-major (row-major [rows][K]);  C fp32 or bf16;  fp32 accumulation in tensor memory.
// The K-major x K-major products of the layer (projection y = T.W, dT = dy.W^T, dx = U.Wp^T), same contract as
// k_gemm_tcgen05 (proj_tcgen05.cu), which this kernel supersedes when its shapes allow:
//   * warp 0, one thread: cp.async.bulk.tensor.2d (SASS UTMALDG) of the [128 x 64] A box and the [BN x 64] B box into a
//     ring of SWIZZLE_128B stages; completion by mbarrier transaction count (no thread ever touches the operands);
//   * warp 1, one thread: four tcgen05.mma (kind::f16, M = 128, N = BN, K = 16) per stage, tcgen05.commit frees the stage;
//     the accumulator is DOUBLE-BUFFERED in tensor memory, so the MMAs of the next tile run while the epilogue drains
//     the previous one;
//   * warps 2..5: epilogue -- tcgen05.ld (32 columns at a time), scale, warp-private swizzled staging tile in shared
//     memory, full 128-byte line stores;
//   * persistent: one CTA per SM walks (m-tile, n-tile, k-split) work items in launch order; split-K (ordered reduction
//     by k_splitk_reduce) when the output tiles alone cannot fill the GPU.  Work items are ordered n-tile fastest, so
//     the column tiles of one row tile run on neighbouring CTAs at the same time and the tall operand is read from HBM
//     once (the repeat hits L2), and row tiles back to front: the tail of the operand the producing kernel has just
//     written is still in L2.
// Tensor maps are encoded on the host (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: the library links only
// libcudart) and passed as __grid_constant__ kernel parameters; out-of-range rows / k are zero-filled by the hardware.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"
#include "tc_common.cuh"
#include "umma.cuh"

namespace se3 {

void splitk_reduce_launch(const float* partials, int splits, int64_t mn, float alpha, float* c, cudaStream_t st);

using namespace umma;

namespace {

constexpr int TBM = 128, TBK = 64;
// warp 0: TMA, warp 1: MMA, then one or two groups of four epilogue warps (one warp per TMEM lane quadrant and group; with two
// groups -- tiles of at least 128 columns -- each group drains half of the accumulator's columns: the bf16-output products
// of the layer (dT = dy W^T, K = Cout) are all epilogue)
constexpr int epi_groups(int bn) { return bn >= 128 ? 2 : 1; }
constexpr int nthr(int bn) { return 64 + 128 * epi_groups(bn); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// MN = false: A [M][K], B [N][K] (K-major operands).  MN = true: A stored [K][M], B stored [K][N] -- the contraction index
// is the ROW index of both operands (dW = T^T dy): every [64 k-rows x 64 columns] TMA box lands as one canonical UMMA
// MN-major SWIZZLE_128B atom (128-byte lines of 64 consecutive columns, one line per k), 64-column atoms 8 KB apart.
template <int BN, bool OUT_BF16, bool MN = false>
__global__ void __launch_bounds__(nthr(BN), 1) k_gemm_tma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                      int M, int N, int K, float alpha, void* __restrict__ Cout, int64_t ldc,
                                                      int m_tiles, int n_tiles, int splits, int kb_per_split,
                                                      int64_t split_stride, int nstages) {
  constexpr uint32_t A_BYTES = TBM * 128, B_BYTES = BN * 128;
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;                  // columns of one accumulator
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
  constexpr int MAXST = 8;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t full[MAXST], empty[MAXST], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_addr(smem_dyn);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  const uint32_t stage_bytes = A_BYTES + B_BYTES;
  const uint32_t epi = ring + nstages * stage_bytes;               // 4 x 4 KB staging tiles of the epilogue warps
  unsigned char* epi_gen = smem_dyn + (epi - raw);
  const int nkb_all = (K + TBK - 1) / TBK;
  const int n_work = m_tiles * n_tiles * splits;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int i = 0; i < nstages; ++i) {
      mbar_init(smem_addr(&full[i]), 1);
      mbar_init(smem_addr(&empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_addr(&acc_full[i]), 1);
      mbar_init(smem_addr(&acc_empty[i]), 4 * epi_groups(BN));
    }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc(smem_addr(&tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();   // everything below touches global memory the previous kernel may have written

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int nt = w % n_tiles, rest = w / n_tiles, mt = m_tiles - 1 - rest % m_tiles, sp = rest / m_tiles;
        const int kb0 = sp * kb_per_split, nkb = min(kb_per_split, nkb_all - kb0);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_addr(&empty[s]), ph ^ 1u);
          const uint32_t a = ring + s * stage_bytes, b = a + A_BYTES;
          mbar_arrive_expect_tx(smem_addr(&full[s]), stage_bytes);
          if (MN) {
#pragma unroll
            for (int at = 0; at < TBM / 64; ++at)
              tma_load_2d(a + at * 8192, &map_a, mt * TBM + at * 64, (kb0 + kb) * TBK, smem_addr(&full[s]));
#pragma unroll
            for (int at = 0; at < BN / 64; ++at)
              tma_load_2d(b + at * 8192, &map_b, nt * BN + at * 64, (kb0 + kb) * TBK, smem_addr(&full[s]));
          } else {
            tma_load_2d(a, &map_a, (kb0 + kb) * TBK, mt * TBM, smem_addr(&full[s]));
            tma_load_2d(b, &map_b, (kb0 + kb) * TBK, nt * BN, smem_addr(&full[s]));
          }
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID = idesc(FMT_BF16, FMT_BF16, TBM, BN, MN, MN);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int sp = (w / n_tiles) / m_tiles;
        const int kb0 = sp * kb_per_split, nkb = min(kb_per_split, nkb_all - kb0);
        const int ab = it & 1;
        mbar_wait(smem_addr(&acc_empty[ab]), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d = tmem + (uint32_t)(ab * ACC_COLS);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_addr(&full[s]), ph);
          tc_fence_after();
          if (MN) {
            // advancing K by 16 = two 8-line k-groups = 2048 bytes (128 descriptor units)
            const uint64_t ad = desc_mnmajor_sw128(ring + s * stage_bytes, 8192), bd = desc_mnmajor_sw128(ring + s * stage_bytes + A_BYTES, 8192);
#pragma unroll
            for (int k4 = 0; k4 < TBK / 16; ++k4) mma_f16(d, ad + (uint64_t)(128 * k4), bd + (uint64_t)(128 * k4), ID, (kb | k4) ? 1u : 0u);
          } else {
            const uint64_t ad = desc_kmajor_sw128(ring + s * stage_bytes), bd = desc_kmajor_sw128(ring + s * stage_bytes + A_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < TBK / 16; ++k4) mma_f16(d, ad + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), ID, (kb | k4) ? 1u : 0u);
          }
          commit(smem_addr(&empty[s]));
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
        commit(smem_addr(&acc_full[ab]));
      }
    }
  } else {
    const int q = warp & 3;                     // TMEM lane quadrant of this warp
    unsigned char* stg = epi_gen + (warp - 2) * 4096;   // [32 rows][128 B], 16-byte chunks XOR-swizzled by the row
    constexpr int CPG = BN / epi_groups(BN);            // accumulator columns of one epilogue group
    const int cbeg = ((warp - 2) >> 2) * CPG;
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int nt = w % n_tiles, rest = w / n_tiles, mt = m_tiles - 1 - rest % m_tiles, sp = rest / m_tiles;
      const int ab = it & 1;
      mbar_wait(smem_addr(&acc_full[ab]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const int row0 = mt * TBM + q * 32, col0 = nt * BN;
      const float sc = split_stride ? 1.0f : alpha;
      constexpr int CW = OUT_BF16 ? 64 : 32;    // columns per staging pass (128 bytes per row)
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + CPG; c0 += CW) {
        uint32_t r[32], r2[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * ACC_COLS + c0), r);
        if (OUT_BF16 && c0 + 32 < BN) tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * ACC_COLS + c0 + 32), r2);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4 p;
          if (OUT_BF16) {
            const uint32_t* src = ch < 4 ? r + ch * 8 : r2 + (ch - 4) * 8;
            p.x = pack_bf16(sc * __uint_as_float(src[0]), sc * __uint_as_float(src[1]));
            p.y = pack_bf16(sc * __uint_as_float(src[2]), sc * __uint_as_float(src[3]));
            p.z = pack_bf16(sc * __uint_as_float(src[4]), sc * __uint_as_float(src[5]));
            p.w = pack_bf16(sc * __uint_as_float(src[6]), sc * __uint_as_float(src[7]));
          } else {
            p.x = __float_as_uint(sc * __uint_as_float(r[4 * ch + 0]));
            p.y = __float_as_uint(sc * __uint_as_float(r[4 * ch + 1]));
            p.z = __float_as_uint(sc * __uint_as_float(r[4 * ch + 2]));
            p.w = __float_as_uint(sc * __uint_as_float(r[4 * ch + 3]));
          }
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((ch ^ (lane & 7)) << 4)) = p;
        }
        __syncwarp();
        // four rows per store instruction, eight 16-byte chunks per row: full 128-byte lines
        constexpr int EPC = OUT_BF16 ? 8 : 4;   // elements per 16-byte chunk
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = k * 4 + (lane >> 3), ch = lane & 7;
          const int grow = row0 + rr, gcol = col0 + c0 + ch * EPC;
          if (grow < M && gcol < N) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
            if (OUT_BF16) {
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(Cout) + (int64_t)grow * ldc + gcol) = v;
            } else {
              float* dst = reinterpret_cast<float*>(Cout) + (split_stride ? (int64_t)sp * split_stride : 0) + (int64_t)grow * ldc + gcol;
              *reinterpret_cast<uint4*>(dst) = v;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&acc_empty[ab]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

// [rows][K] bf16 row-major (ld elements between rows) -> tensor map with a [box_rows x 64] SWIZZLE_128B box
bool make_map(CUtensorMap* m, const __nv_bfloat16* base, int64_t rows, int64_t k, int64_t ld, int box_rows) {
  EncodeFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// operand stored [k_rows][cols] bf16 row-major -> tensor map with a [64 k-rows x 64 columns] SWIZZLE_128B box
bool make_map_mn(CUtensorMap* m, const __nv_bfloat16* base, int64_t k_rows, int64_t cols, int64_t ld) {
  EncodeFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)k_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)TBK};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// C[M,N] = alpha * A^T B, A stored [K][M], B stored [K][N], fp32 output; split-K over the rows with ordered reduction
template <int BN>
int launch_tma_mn_cfg(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                      int64_t ldb, float* c, int64_t ldc, int splits, float* partials, cudaStream_t st) {
  CUtensorMap ma, mb;
  if (!make_map_mn(&ma, a, k, m, lda) || !make_map_mn(&mb, b, k, n, ldb)) {
    set_error("launch_gemm_tma_mn: cuTensorMapEncodeTiled failed");
    return SE3_ECUDA;
  }
  auto kern = k_gemm_tma<BN, false, true>;
  const int nkb = (int)((k + TBK - 1) / TBK);
  if (partials == nullptr || splits < 1 || ldc != n) splits = 1;
  int per = (nkb + splits - 1) / splits;
  splits = (nkb + per - 1) / per;
  const int m_tiles = (int)((m + TBM - 1) / TBM), n_tiles = (int)((n + BN - 1) / BN);
  const size_t stage = (size_t)TBM * 128 + (size_t)BN * 128;
  int nstages = (int)std::min<size_t>(8, (226 * 1024 - 4 * epi_groups(BN) * 4096) / stage);   // 227 KB per CTA: ring + staging + barriers
  if (nstages < 2) nstages = 2;
  const size_t smem = (size_t)nstages * stage + 4 * epi_groups(BN) * 4096 + 1024;
  SE3_SMEM_ONCE(kern, smem);
  const int64_t work = (int64_t)m_tiles * n_tiles * splits;
  const int grid = (int)std::min<int64_t>(work, num_sms());
  if (splits > 1) {
    SE3_CUDA(launch_pdl(kern, dim3(grid), dim3(nthr(BN)), smem, st, ma, mb, (int)m, (int)n, (int)k, 1.0f, (void*)partials, n, m_tiles,
                        n_tiles, splits, per, m * n, nstages));
    SE3_LAUNCH_CHECK();
    splitk_reduce_launch(partials, splits, m * n, alpha, c, st);
    SE3_LAUNCH_CHECK();
  } else {
    SE3_CUDA(launch_pdl(kern, dim3(grid), dim3(nthr(BN)), smem, st, ma, mb, (int)m, (int)n, (int)k, alpha, (void*)c, ldc, m_tiles,
                        n_tiles, 1, nkb, (int64_t)0, nstages));
    SE3_LAUNCH_CHECK();
  }
  return SE3_OK;
}

template <int BN, bool OB>
int launch_tma_cfg(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                   int64_t ldb, void* c, int64_t ldc, int splits, float* partials, cudaStream_t st) {
  CUtensorMap ma, mb;
  if (!make_map(&ma, a, m, k, lda, TBM) || !make_map(&mb, b, n, k, ldb, BN)) {
    set_error("launch_gemm_tma: cuTensorMapEncodeTiled failed");
    return SE3_ECUDA;
  }
  auto kern = k_gemm_tma<BN, OB>;
  const int nkb = (int)((k + TBK - 1) / TBK);
  if (OB || partials == nullptr || splits < 1 || ldc != n) splits = 1;
  int per = (nkb + splits - 1) / splits;
  splits = (nkb + per - 1) / per;
  const int m_tiles = (int)((m + TBM - 1) / TBM), n_tiles = (int)((n + BN - 1) / BN);
  const size_t stage = (size_t)TBM * 128 + (size_t)BN * 128;
  int nstages = (int)std::min<size_t>(8, (226 * 1024 - 4 * epi_groups(BN) * 4096) / stage);   // 227 KB per CTA: ring + staging + barriers
  if (nstages < 2) nstages = 2;
  const size_t smem = (size_t)nstages * stage + 4 * epi_groups(BN) * 4096 + 1024;
  SE3_SMEM_ONCE(kern, smem);
  const int64_t work = (int64_t)m_tiles * n_tiles * splits;
  const int grid = (int)std::min<int64_t>(work, num_sms());
  if (splits > 1) {
    SE3_CUDA(launch_pdl(kern, dim3(grid), dim3(nthr(BN)), smem, st, ma, mb, (int)m, (int)n, (int)k, 1.0f, (void*)partials, n, m_tiles,
                        n_tiles, splits, per, m * n, nstages));
    SE3_LAUNCH_CHECK();
    splitk_reduce_launch(partials, splits, m * n, alpha, reinterpret_cast<float*>(c), st);
    SE3_LAUNCH_CHECK();
  } else {
    SE3_CUDA(launch_pdl(kern, dim3(grid), dim3(nthr(BN)), smem, st, ma, mb, (int)m, (int)n, (int)k, alpha, c, ldc, m_tiles, n_tiles,
                        1, nkb, (int64_t)0, nstages));
    SE3_LAUNCH_CHECK();
  }
  return SE3_OK;
}

}  // namespace

// shapes the TMA kernel takes: full 16-byte chunks in every output row, 16-byte aligned operand rows
bool tma_gemm_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc, bool out_bf16, const void* a,
                        const void* b, const void* c) {
  static const bool off = getenv("SE3_GEMM_TMA") && getenv("SE3_GEMM_TMA")[0] == '0';
  if (off || encode_fn() == nullptr) return false;
  const int epc = out_bf16 ? 8 : 4;
  return m >= 1 && n >= 16 && (n % 16) == 0 && (k % 8) == 0 && (lda % 8) == 0 && (ldb % 8) == 0 && (ldc % epc) == 0 &&
         (n % epc) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(c) & 15) == 0 && m < ((int64_t)1 << 31) && k < ((int64_t)1 << 31);
}

int launch_gemm_tma(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                    int64_t ldb, void* c, int64_t ldc, bool out_bf16, int splits, float* partials, cudaStream_t st) {
  const int bn = n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)));
#define SE3_TMA_CASE(BN_)                                                                                          \
  case BN_:                                                                                                         \
    return out_bf16 ? launch_tma_cfg<BN_, true>(m, n, k, alpha, a, lda, b, ldb, c, ldc, 1, nullptr, st)             \
                    : launch_tma_cfg<BN_, false>(m, n, k, alpha, a, lda, b, ldb, c, ldc, splits, partials, st);
  switch (bn) {
    SE3_TMA_CASE(16)
    SE3_TMA_CASE(32)
    SE3_TMA_CASE(64)
    SE3_TMA_CASE(128)
    SE3_TMA_CASE(256)
  }
#undef SE3_TMA_CASE
  return SE3_EINVAL;
}

// dW = T^T dy on the TMA kernel: 16-byte aligned operand rows, whole 16-byte chunks per output row
bool tma_gemm_mn_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc, const void* a, const void* b,
                           const void* c) {
  static const bool off = getenv("SE3_GEMM_TMA") && getenv("SE3_GEMM_TMA")[0] == '0';
  if (off || encode_fn() == nullptr) return false;
  return m >= 1 && n >= 8 && (n % 8) == 0 && (m % 8) == 0 && (lda % 8) == 0 && (ldb % 8) == 0 && (ldc % 4) == 0 && k >= 1 &&
         (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(c) & 15) == 0 && m < ((int64_t)1 << 31) && k < ((int64_t)1 << 31);
}

int launch_gemm_tma_mn(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                       int64_t ldb, float* c, int64_t ldc, int splits, float* partials, cudaStream_t st) {
  if (n <= 64) return launch_tma_mn_cfg<64>(m, n, k, alpha, a, lda, b, ldb, c, ldc, splits, partials, st);
  if (n <= 128) return launch_tma_mn_cfg<128>(m, n, k, alpha, a, lda, b, ldb, c, ldc, splits, partials, st);
  return launch_tma_mn_cfg<256>(m, n, k, alpha, a, lda, b, ldb, c, ldc, splits, partials, st);
}

}  // namespace se3
