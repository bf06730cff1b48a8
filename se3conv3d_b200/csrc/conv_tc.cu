// bf16 tensor-core path (precision 1) of the fused PNEConvLayerRotEquiv.
//
//   k_agg_tc   : one warp per (row point, frame group, channel block) item, persistent grid.  Per chunk of 32 expanded
//                neighbours (edge x in-frame): records and bf16 feature rows arrive by cp.async (double buffered), lanes
//                build the 9-vector g per (neighbour, row frame) into 32-byte f16 rows, the basis affine runs as ONE f16
//                mma.m16n8k16 per 8 neighbours x 16 basis functions and h = act(g.W9 + b) lands DIRECTLY in the B-fragment
//                layout of the aggregation mma (GELU in packed bf16x2), T[c,k] += x[n,c] h[n,k] with mma.sync m16n8k16
//                (bf16 in, fp32 accumulate).  g, h and the expanded neighbour list never reach HBM.
//                TR=false: rows = output points (forward).  TR=true: rows = input points over the transposed CSR with dy
//                as the gathered feature (atomic-free data gradient).
//   k_agg_wide : more than 64 channels: one warp per (row point, row frame) walks all 64-channel blocks, the basis
//                fragments of the row stay in shared memory after the first block.
//   k_edge_row_tc / k_edge_tc : by output row, gradient of proj_axes_ / proj_biases_: dH^T = dT^T X^T (mma), times
//                act'(pre) with pre recomputed in the accumulator layout (f16 mma, packed bf16x2 GELU'), then
//                [k x n].[n x 10] (mma) against the bf16 geometry (+ a ones column for the bias); per-CTA partials,
//                ordered reduce.  DX variant + k_dx_segsum: the opt-in merged backward pass (SE3_BWD_MERGED=1).
//   k_gemm_bf16: generic bf16 tensor-core GEMM (cp.async 3-stage pipeline, ldmatrix, mma.sync), the fallback of the
//                projection and its three backward products; deterministic split-K.
// The tcgen05 / TMEM GEMMs live in proj_tma.cu (persistent, TMA-fed, warp-specialised: the default) and
// proj_tcgen05.cu (cp.async operand loads); conv_tc_fwd / conv_tc_bwd at the end of this file orchestrate a layer call.
#include <stdlib.h>
#include <algorithm>
#include "conv_simt.cuh"
#include "conv_fused.cuh"
#include "tc_common.cuh"

namespace se3 {

void splitk_reduce_launch(const float* partials, int splits, int64_t mn, float alpha, float* c, cudaStream_t st);
// proj_tcgen05.cu
bool tcgen05_gemm_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb);
int tcgen05_gemm_splits(int64_t m, int64_t n, int64_t k);
int launch_gemm_tcgen05(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda,
                        const __nv_bfloat16* b, int64_t ldb, void* c, int64_t ldc, bool out_bf16, int splits,
                        float* partials, cudaStream_t st);
bool tcgen05_gemm_mn_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb);
// proj_tma.cu
bool tma_gemm_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc, bool out_bf16, const void* a,
                        const void* b, const void* c);
int launch_gemm_tma(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                    int64_t ldb, void* c, int64_t ldc, bool out_bf16, int splits, float* partials, cudaStream_t st);
bool tma_gemm_mn_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc, const void* a, const void* b,
                           const void* c);
int launch_gemm_tma_mn(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                       int64_t ldb, float* c, int64_t ldc, int splits, float* partials, cudaStream_t st);
int launch_gemm_tcgen05_mn(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda,
                           const __nv_bfloat16* b, int64_t ldb, float* c, int64_t ldc, int splits, float* partials,
                           cudaStream_t st);

struct TcAggArgs {
  const int* row_ends;
  const int* nbr;
  const float* rec_row;  // [n_rows * FR, 12] records of the row side
  const float* rec_g;    // [n_g * f_g, 12] records of the gathered side
  int f_g;
  const __nv_bfloat16* feat;  // [n_g * f_g, cs] bf16 rows
  int c;                      // channels
  int cs;                     // row stride of feat (multiple of 8, >= c)
  const float* w9;
  const float* bias;
  float norm;
  int act;
  __nv_bfloat16* out;
  int64_t n_rows;
  int f_row;  // frames per row point (the kernel's FR frames per item may be a divisor of it: frame groups)
};

// geometry of one (row frame, gathered frame) pair for this lane's neighbour
template <bool TR>
__device__ __forceinline__ void geometry9(const float (&Frow)[9], const float (&Fq)[9], float dx, float dy, float dz,
                                          float (&g)[9]) {
  const float(&Ro)[9] = TR ? Fq : Frow;
  const float(&Ri)[9] = TR ? Frow : Fq;
#pragma unroll
  for (int c = 0; c < 3; ++c) g[c] = dx * Ro[c] + dy * Ro[3 + c] + dz * Ro[6 + c];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 3; ++n) g[3 + 3 * m + n] = Ro[m] * Ri[n] + Ro[3 + m] * Ri[3 + n] + Ro[6 + m] * Ri[6 + n];
}

template <int ACT>
__device__ __forceinline__ float act_rt(float x, int act) {
  if (ACT >= 0) return act_fast<ACT>(x);
  switch (act) {
    case 1: return act_fast<1>(x);
    case 2: return act_fast<2>(x);
    case 3: return act_fast<3>(x);
    default: return x;
  }
}
template <int ACT>
__device__ __forceinline__ float act_grad_rt(float x, int act) {
  if (ACT >= 0) return act_grad_fast<ACT>(x);
  switch (act) {
    case 1: return act_grad_fast<1>(x);
    case 2: return act_grad_fast<2>(x);
    case 3: return act_grad_fast<3>(x);
    default: return 1.0f;
  }
}

// pair forms: (z0, z1) -> bf16x2 of act(z), and bf16x2 of d * act'(z).  The compiled-in GELU runs packed
// (tc_common.cuh), everything else in fp32 with one rounding at the end.
template <int ACT>
__device__ __forceinline__ uint32_t act_pair(float z0, float z1, int act) {
#if SE3_PACKED_ACT
  if (ACT == 2) return gelu_half_arg_bf2(pack_bf16(z0, z1));
#endif
  return pack_bf16(act_rt<ACT>(z0, act), act_rt<ACT>(z1, act));
}
template <int ACT>
__device__ __forceinline__ uint32_t act_grad_pair(float d0, float d1, float z0, float z1, int act) {
#if SE3_PACKED_ACT
  if (ACT == 2) return bf2_mul(pack_bf16(d0, d1), gelu_grad_half_arg_bf2(pack_bf16(z0, z1)));
#endif
  return pack_bf16(d0 * act_grad_rt<ACT>(z0, act), d1 * act_grad_rt<ACT>(z1, act));
}

// d * act'(z) as above, and act(z) for the same pair (merged backward pass): one tanh serves both
template <int ACT>
__device__ __forceinline__ uint32_t act_both_pair(float d0, float d1, float z0, float z1, int act, uint32_t& h) {
#if SE3_PACKED_ACT
  if (ACT == 2) {
    uint32_t g;
    gelu_both_half_arg_bf2(pack_bf16(z0, z1), h, g);
    return bf2_mul(pack_bf16(d0, d1), g);
  }
#endif
  h = pack_bf16(act_rt<ACT>(z0, act), act_rt<ACT>(z1, act));
  return pack_bf16(d0 * act_grad_rt<ACT>(z0, act), d1 * act_grad_rt<ACT>(z1, act));
}

// n -> (edge, frame) for f frames per gathered point (f in 1..4, warp-uniform)
__device__ __forceinline__ void split_nf(int n, int f, int& e, int& fg) {
  if (f == 1) {
    e = n; fg = 0;
  } else if (f == 2) {
    e = n >> 1; fg = n & 1;
  } else if (f == 4) {
    e = n >> 2; fg = n & 3;
  } else {
    e = n / 3; fg = n - 3 * e;
  }
}

// 16-byte asynchronous copy through L1 (neighbouring rows share gathered lines), zero fill when !pred
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// Gather of one 32-neighbour chunk, fully asynchronous (no registers in flight):
//   * this lane's neighbour record (48 B) -> Rg[lane][12]
//   * the bf16 feature rows of all 32 neighbours, channels [c0, c0+CB) -> Xs[32][CB+8]; CB/8 lanes per row.
// gidx < 0 marks a padding lane (zero filled).  feat rows are `cs` bf16 apart (cs % 8 == 0).
template <int CB>
__device__ __forceinline__ void gather_chunk_async(const float* __restrict__ rec_g, const __nv_bfloat16* __restrict__ feat,
                                                   int cs, int c0, int gidx, int lane, uint32_t rg_s, uint32_t xs_s) {
  constexpr int XS = CB + 8;
  {
    const bool ok = gidx >= 0;
    const float* src = rec_g + (ok ? (uint32_t)gidx * 12u : 0u);
    const uint32_t dst = rg_s + lane * 48;
    cp_async16_ca(dst, src, ok);
    cp_async16_ca(dst + 16, src + 4, ok);
    cp_async16_ca(dst + 32, src + 8, ok);
  }
  constexpr int LPR = CB / 8;    // lanes per row (16 B = 8 bf16 each)
  constexpr int RPI = 32 / LPR;  // rows per iteration
  // CB = 32 (80-byte rows, four lanes each): a quarter-warp takes rows q and q + 4 of the eight rows of an iteration --
  // their 16-byte chunks fall into eight different bank groups (rows q, q + 1 collide in one: 5 chunks apart)
  const int col = lane % LPR, r0 = CB == 32 ? (lane >> 3) + 4 * ((lane >> 2) & 1) : lane / LPR;
  const int ch = c0 + col * 8;
  const bool chok = ch < cs;
#pragma unroll
  for (int it = 0; it < LPR; ++it) {
    const int row = it * RPI + r0;
    const int src = __shfl_sync(0xffffffffu, gidx, row);
    const bool ok = src >= 0 && chok;
    cp_async16_ca(xs_s + (row * XS + col * 8) * 2, feat + (ok ? (uint32_t)(src * cs + ch) : 0u), ok);
  }
}

// This lane's gathered index of chunk `base` of a row with `nt` (edge x frame) entries: -1 when padding.
// idx holds the row's first 32 neighbour ids (lane e <-> edge e); longer rows fall back to a direct load.
__device__ __forceinline__ int chunk_gidx(const int* __restrict__ nbr, int lo, int nt, int idx, int f_g, int base, int lane) {
  const int n = base + lane;
  const bool valid = n < nt;
  int e, fg;
  split_nf(valid ? n : 0, f_g, e, fg);
  int q;
  if (base + 31 < 32 * f_g) {  // warp-uniform: every edge of this chunk is among the preloaded 32
    q = __shfl_sync(0xffffffffu, idx, e);
  } else {
    q = __ldg(nbr + lo + e);
  }
  return valid ? q * f_g + fg : -1;
}

constexpr int AGG_WARPS = 4;  // warps per CTA of the aggregation / edge kernels
constexpr int GROW = 32;      // bytes of a geometry row in shared memory: 16 f16 slots (9 components, 1, zeros)

// A fragments of the basis affine on the tensor cores (f16 m16n8k16, fp32 accumulate):
//   pre^T[k, n] = sum_d Wext[d, k] G[n, d],  Wext = [proj_axes_ (9 rows); proj_biases_ (row 9); 0].
// f16 carries the 11 significant bits tf32 does (values are O(1): rotation entries, offsets scaled by 1 / radius,
// trained axes), and one k = 16 instruction replaces a dependent pair of tf32 k = 8 ones.
// The 16 k-slots are permuted so that lane (g, t) needs the four CONSECUTIVE components 4t..4t+3 of a geometry row
// (one 64-bit shared load): slots (2t, 2t+1) <-> components (4t, 4t+1), slots (2t+8, 2t+9) <-> (4t+2, 4t+3).
// Components 12..15 are zero, so lanes t = 3 carry zero fragments.  `scale` (act_pre_scale) is folded in.
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void mma_f16_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}
__device__ __forceinline__ void load_w9_frags(const float* __restrict__ w9, const float* __restrict__ bias, int g, int t,
                                              float scale, uint32_t (&aw)[2][4]) {
  auto wext = [&](int d, int k) -> float {
    return scale * (d < 9 ? __ldg(w9 + d * 32 + k) : (d == 9 ? __ldg(bias + k) : 0.0f));
  };
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int d0 = 4 * t, k0 = 16 * m + g;
    aw[m][0] = pack_f16(wext(d0, k0), wext(d0 + 1, k0));
    aw[m][1] = pack_f16(wext(d0, k0 + 8), wext(d0 + 1, k0 + 8));
    aw[m][2] = pack_f16(wext(d0 + 2, k0), wext(d0 + 3, k0));
    aw[m][3] = pack_f16(wext(d0 + 2, k0 + 8), wext(d0 + 3, k0 + 8));
  }
}

// pre^T tile of one 8-neighbour group: d[m][0..3] = pre[k = 16m+g (+8 for 2,3)][n = 2t (+1 for 1,3)].
// grow_s = shared-window address of the geometry row of neighbour g of the group, plus 8 * min(t, 2): lanes t = 3
// own the slots of components 12..15, whose A fragments are zero -- they re-read the (finite) components 8..11 of
// the lanes t = 2 (a broadcast) instead of branching around the load.
__device__ __forceinline__ void basis_pre(const uint32_t (&aw)[2][4], uint32_t grow_s, float (&d)[2][4]) {
  uint32_t b0, b1;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(grow_s) : "memory");
#pragma unroll
  for (int m = 0; m < 2; ++m) mma_f16_zero(d[m], aw[m], b0, b1);
}

// geometry row of this lane's neighbour: 9 components, the bias input, zeros (f16)
__device__ __forceinline__ void store_geometry_row(uint32_t row_s, int row, const float (&gg)[9], float one) {
  // rows are 32 bytes, one row per lane: the two 16-byte halves of rows 4..7 (mod 8) are swapped, so that the eight rows
  // of a quarter-warp land in eight different 16-byte bank groups (plain 32-byte rows: a two-way conflict per store)
  const uint32_t sw = ((uint32_t)row & 4u) << 2;   // 16 for rows 4..7 (mod 8)
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_s + sw), "r"(pack_f16(gg[0], gg[1])),
               "r"(pack_f16(gg[2], gg[3])), "r"(pack_f16(gg[4], gg[5])), "r"(pack_f16(gg[6], gg[7]))
               : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(row_s + 16u - sw), "r"(pack_f16(gg[8], one)), "r"(0u) : "memory");
}
// byte offset, inside its group of eight geometry rows, of the 8 bytes lane (g, t) feeds to the basis mma: components
// 4 min(t, 2) .. + 3 of row g, with the half swap of store_geometry_row
__device__ __forceinline__ uint32_t geom_lane_off(int g, int t) {
  const int tt = min(t, 2);
  return (uint32_t)(g * GROW + ((((tt >> 1) ^ (g >> 2)) & 1) << 4) + ((tt & 1) << 3));
}

// This lane's gathered neighbour of a chunk from its staged record: scaled offset and frame.
template <bool TR>
__device__ __forceinline__ void unpack_neighbour(const float* rg, float prx, float pry, float prz, float norm, float& dx,
                                                 float& dy, float& dz, float (&Fq)[9]) {
  const float4* rp = reinterpret_cast<const float4*>(rg);
  const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
  dx = (TR ? (prx - r0.x) : (r0.x - prx)) * norm;
  dy = (TR ? (pry - r0.y) : (r0.y - pry)) * norm;
  dz = (TR ? (prz - r0.z) : (r0.z - prz)) * norm;
  Fq[0] = r0.w; Fq[1] = r1.x; Fq[2] = r1.y; Fq[3] = r1.z; Fq[4] = r1.w;
  Fq[5] = r2.x; Fq[6] = r2.y; Fq[7] = r2.z; Fq[8] = r2.w;
}

__device__ __forceinline__ void load_row_frame(const float* Rs, int f, float (&Frow)[9]) {
  const float4* r = reinterpret_cast<const float4*>(Rs + f * 12);
  const float4 r0 = r[0], r1 = r[1], r2 = r[2];
  Frow[0] = r0.w; Frow[1] = r1.x; Frow[2] = r1.y; Frow[3] = r1.z; Frow[4] = r1.w;
  Frow[5] = r2.x; Frow[6] = r2.y; Frow[7] = r2.z; Frow[8] = r2.w;
}

// One aggregation k-step (16 neighbours, NG = 1 or 2 groups of 8) for all FR row frames, straight-line code:
// basis h = act(pre) is produced directly as the B fragments of the aggregation mma -- n-tile j (k = 8j + g),
// b0 = neighbours 2t,2t+1, b1 = neighbours 2t+8,2t+9 -- and T[c,k] += x[n,c] h[n,k].
template <int CB, int FR, int NG, int ACT>
__device__ __forceinline__ void agg_kstep(float (&acc)[FR][CB / 16][4][4], const uint32_t (&aw)[2][4], uint32_t gs_s,
                                          int g_frame_bytes, const __nv_bfloat16* Xs, int ks, int lane, int act) {
  constexpr int XS = CB + 8, MT = CB / 16;
  const int g = lane >> 2, t = lane & 3, mid = lane >> 3, mr = lane & 7;
  // A fragments (x^T): [m = channel][k = neighbour], from Xs[n][c] through ldmatrix.trans
  uint32_t af[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
    ldmatrix_x4_trans(af[m][0], af[m][1], af[m][2], af[m][3],
                      smem_u32(Xs + (ks * 16 + (mid >> 1) * 8 + mr) * XS + m * 16 + (mid & 1) * 8));
#pragma unroll
  for (int f = 0; f < FR; ++f) {
    uint32_t hb[4][2];
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
      if (hq < NG) {
        float d[2][4];
        basis_pre(aw, gs_s + f * g_frame_bytes + (ks * 16 + hq * 8) * GROW + geom_lane_off(g, t), d);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          hb[2 * m][hq] = act_pair<ACT>(d[m][0], d[m][1], act);
          hb[2 * m + 1][hq] = act_pair<ACT>(d[m][2], d[m][3], act);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) hb[j][hq] = 0u;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int m = 0; m < MT; ++m) mma_bf16(acc[f][m][j], af[m], hb[j][0], hb[j][1]);
  }
}

template <int CB, int FR>
struct AggSmem {
  static constexpr int XS = CB + 8;
  static constexpr int X_BYTES = 32 * XS * 2;         // one feature buffer (bf16)
  static constexpr int RG_BYTES = 32 * 48;            // one record buffer
  // per row frame: geometry rows [32][GROW B], reused as the T staging tile [CB][64 B], whichever is larger
  static constexpr int G_FRAME_BYTES = (32 * GROW > CB * 64) ? 32 * GROW : CB * 64;
  static constexpr int G_BYTES = FR * G_FRAME_BYTES;
  static constexpr int RS_BYTES = ((FR * 48 + 63) / 64) * 64;  // one row-record buffer
  static constexpr int OFF_RG = 2 * X_BYTES;
  static constexpr int OFF_G = OFF_RG + 2 * RG_BYTES;
  static constexpr int OFF_RS = OFF_G + G_BYTES;
  static constexpr int WARP_BYTES = OFF_RS + 2 * RS_BYTES;
};

#ifndef SE3_AGG_MIN_BLOCKS
#define SE3_AGG_MIN_BLOCKS 3
#endif
#ifndef SE3_AGG_WAVES
#define SE3_AGG_WAVES 1
#endif

// One warp per (row point, channel block): the FR row frames share the gathered records and the staged
// feature rows; their [CB x 32] accumulators live in registers.
//
// Software pipeline (per warp, no block-level synchronisation): while chunk i is computed, the records and
// feature rows of chunk i+1 -- the next chunk of the same row, or the first chunk of the warp's next row --
// are in flight as cp.async copies into the other shared-memory buffer; the neighbour ids of the next row
// and the CSR bounds of the row after it are register prefetches issued one row earlier.  No gather latency
// sits on the critical path after the first chunk.
// FG: gathered frames per point known at compile time (2: the entry -> (edge, frame) split is a shift / mask), 0: a.f_g
template <int CB, int FR, bool TR, int ACT, int FG = 0>
__global__ void __launch_bounds__(AGG_WARPS * 32, SE3_AGG_MIN_BLOCKS) k_agg_tc(const TcAggArgs a, const int ncb) {
  using SM = AggSmem<CB, FR>;
  constexpr int MT = CB / 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int f_g = FG ? FG : a.f_g;
  unsigned char* wbase = smem_raw + wib * SM::WARP_BYTES;
  const uint32_t wbase_s = smem_u32(wbase);
  unsigned char* Gs = wbase + SM::OFF_G;        // [FR][G_FRAME_BYTES]
  const uint32_t gs_s = wbase_s + SM::OFF_G;
  const int g = lane >> 2, t = lane & 3;
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  // an item = (row point, frame group, channel block); nfg = f_row / FR frame groups of FR row frames each.  The
  // basis is evaluated once per (row frame, gathered frame) pair whatever the split: frame groups only repeat the
  // gather, channel blocks repeat the basis too -- so many frames are split into groups first.
  const int nfg = a.f_row / FR, per_row = ncb * nfg;
  const int total = (int)a.n_rows * per_row;
  int item0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  pdl_wait();
  pdl_trigger();
  if (item0 >= total) return;
  uint32_t aw[2][4];
  load_w9_frags(a.w9, a.bias, g, t, act_pre_scale(a.act), aw);

  // c0 carries the channel offset in its low 16 bits and the first row frame of the item above them
  auto row_of = [&](int item, int& rp, int& c0) {
    rp = item;
    c0 = 0;
    if (per_row > 1) {
      rp = item / per_row;
      const int sub = item - rp * per_row, fgi = sub / ncb;
      c0 = ((sub - fgi * ncb) * CB) | ((fgi * FR) << 16);
    }
  };
  auto issue = [&](int rp, int c0, int lo, int nt, int idx, int base, int buf, int rb, bool with_row) {
    const int gidx = chunk_gidx(a.nbr, lo, nt, idx, f_g, base, lane);
    gather_chunk_async<CB>(a.rec_g, a.feat, a.cs, c0 & 0xffff, gidx, lane, wbase_s + SM::OFF_RG + buf * SM::RG_BYTES,
                           wbase_s + buf * SM::X_BYTES);
    if (with_row && lane < FR * 3)
      cp_async16_ca(wbase_s + SM::OFF_RS + rb * SM::RS_BYTES + lane * 16,
                    reinterpret_cast<const float4*>(a.rec_row) + ((int64_t)rp * a.f_row + (c0 >> 16)) * 3 + lane, true);
  };

  // ---- prologue: row 0 bounds + ids (synchronous, once), row 1 bounds (synchronous) + ids (in flight),
  //      row 2 bounds (in flight)
  int rp0, c00;
  row_of(item0, rp0, c00);
  int lo0 = rp0 > 0 ? __ldg(a.row_ends + rp0 - 1) : 0;
  int ne0 = __ldg(a.row_ends + rp0) - lo0;
  int idx0 = lane < ne0 ? __ldg(a.nbr + lo0 + lane) : 0;
  int item1 = item0 + nwarps, rp1 = 0, c01 = 0, lo1 = 0, ne1 = 0, idx1 = 0;
  if (item1 < total) {
    row_of(item1, rp1, c01);
    lo1 = rp1 > 0 ? __ldg(a.row_ends + rp1 - 1) : 0;
    ne1 = __ldg(a.row_ends + rp1) - lo1;
    idx1 = lane < ne1 ? __ldg(a.nbr + lo1 + lane) : 0;
  }
  int item2 = item1 + nwarps, e2a = 0, e2b = 0;
  if (item2 < total) {
    int rp2, c02;
    row_of(item2, rp2, c02);
    e2a = rp2 > 0 ? __ldg(a.row_ends + rp2 - 1) : 0;
    e2b = __ldg(a.row_ends + rp2);
  }
  int buf = 0, rb = 0;
  issue(rp0, c00, lo0, ne0 * f_g, idx0, 0, 0, 0, true);
  cp_async_commit();

  for (;;) {
    const int nt0 = ne0 * f_g;
    float acc[FR][MT][4][4];
#pragma unroll
    for (int f = 0; f < FR; ++f)
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[f][m][j][i] = 0.0f;
    int base = 0;
    do {
      // ---- chunk i+1 in flight
      if (base + 32 < nt0) {
        issue(rp0, c00, lo0, nt0, idx0, base + 32, buf ^ 1, rb, false);
      } else if (item1 < total) {
        issue(rp1, c01, lo1, ne1 * f_g, idx1, 0, buf ^ 1, rb ^ 1, true);
      }
      cp_async_commit();
      cp_async_wait<1>();  // chunk i has landed
      __syncwarp();
      const float* Rs = reinterpret_cast<const float*>(wbase + SM::OFF_RS + rb * SM::RS_BYTES);
      {
        const float* rg = reinterpret_cast<const float*>(wbase + SM::OFF_RG + buf * SM::RG_BYTES) + lane * 12;
        float dx, dy, dz, Fq[9];
        unpack_neighbour<TR>(rg, Rs[0], Rs[1], Rs[2], a.norm, dx, dy, dz, Fq);
#pragma unroll
        for (int f = 0; f < FR; ++f) {
          float Frow[9], gg[9];
          load_row_frame(Rs, f, Frow);
          geometry9<TR>(Frow, Fq, dx, dy, dz, gg);
          store_geometry_row(gs_s + f * SM::G_FRAME_BYTES + lane * GROW, lane, gg, 1.0f);
        }
      }
      __syncwarp();
      const __nv_bfloat16* Xs = reinterpret_cast<const __nv_bfloat16*>(wbase + buf * SM::X_BYTES);
      const int n_here = min(32, nt0 - base);  // warp-uniform: 8-neighbour groups beyond it are skipped
#pragma unroll 1
      for (int ks = 0; ks < 2; ++ks) {
        if (ks * 16 >= n_here) break;
        if (ks * 16 + 8 < n_here)
          agg_kstep<CB, FR, 2, ACT>(acc, aw, gs_s, SM::G_FRAME_BYTES, Xs, ks, lane, a.act);
        else
          agg_kstep<CB, FR, 1, ACT>(acc, aw, gs_s, SM::G_FRAME_BYTES, Xs, ks, lane, a.act);
      }
      __syncwarp();
      buf ^= 1;
      base += 32;
    } while (base < nt0);
    // ---- epilogue: accumulators -> bf16 tile [CB][32] in shared memory (16-byte chunks XOR-swizzled by the
    // row pair, conflict free both ways) -> 128-bit coalesced stores of the contiguous [CB x 32] block of T
#pragma unroll
    for (int f = 0; f < FR; ++f) {
      uint32_t* ts = reinterpret_cast<uint32_t*>(Gs + f * SM::G_FRAME_BYTES);  // [CB][16 words]
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r0 = m * 16 + g, r1 = r0 + 8;
          ts[r0 * 16 + ((j ^ ((r0 >> 1) & 3)) << 2) + t] = pack_bf16(acc[f][m][j][0], acc[f][m][j][1]);
          ts[r1 * 16 + ((j ^ ((r1 >> 1) & 3)) << 2) + t] = pack_bf16(acc[f][m][j][2], acc[f][m][j][3]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int f = 0; f < FR; ++f) {
      const uint4* ts = reinterpret_cast<const uint4*>(Gs + f * SM::G_FRAME_BYTES);
      const int ch0 = c00 & 0xffff;
      __nv_bfloat16* o = a.out + ((int64_t)rp0 * a.f_row + (c00 >> 16) + f) * (int64_t)a.c * 32 + (int64_t)ch0 * 32;
#pragma unroll
      for (int it = 0; it < CB / 8; ++it) {
        const int row = it * 8 + (lane >> 2), part = lane & 3;
        if (ch0 + row < a.c) reinterpret_cast<uint4*>(o + row * 32)[part] = ts[row * 4 + (part ^ ((row >> 1) & 3))];
      }
    }
    __syncwarp();
    // ---- rotate the row pipeline
    if (item1 >= total) break;
    item0 = item1; rp0 = rp1; c00 = c01; lo0 = lo1; ne0 = ne1; idx0 = idx1;
    rb ^= 1;
    item1 = item2;
    if (item1 < total) {
      row_of(item1, rp1, c01);
      lo1 = e2a;
      ne1 = e2b - e2a;
      idx1 = lane < ne1 ? __ldg(a.nbr + lo1 + lane) : 0;
    }
    item2 = item1 + nwarps;
    if (item2 < total) {
      int rp2, c02;
      row_of(item2, rp2, c02);
      e2a = rp2 > 0 ? __ldg(a.row_ends + rp2 - 1) : 0;
      e2b = __ldg(a.row_ends + rp2);
    }
  }
  cp_async_wait<0>();
}

template <int CB, int FR, bool TR>
static int launch_agg_cfg(const TcAggArgs& a, int64_t n_g, cudaStream_t st) {
  const size_t smem = AGG_WARPS * AggSmem<CB, FR>::WARP_BYTES;
  const int ncb = (a.c + CB - 1) / CB;
  const int64_t warps = a.n_rows * ncb * (a.f_row / FR);
  if (warps >= (int64_t)1 << 30 || n_g * a.f_g * a.cs >= (int64_t)1 << 31 || (a.cs & 7) || a.cs < a.c) {
    set_error("launch_agg_tc: problem too large for 32-bit row offsets");
    return SE3_EINVAL;
  }
  int64_t blocks = (warps + AGG_WARPS - 1) / AGG_WARPS;
  const int64_t cap = (int64_t)num_sms() * SE3_AGG_MIN_BLOCKS * SE3_AGG_WAVES;  // persistent: the row pipeline
  if (blocks > cap) blocks = cap;                                               // needs many rows per warp
  if (blocks < 1) blocks = 1;
  ProfScope prof(TR ? 1 : 0, st);
  if (a.act == 2 && a.f_g == 2) {
    auto kern = k_agg_tc<CB, FR, TR, 2, 2>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(AGG_WARPS * 32), smem, st, a, ncb));
  } else if (a.act == 2) {
    auto kern = k_agg_tc<CB, FR, TR, 2>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(AGG_WARPS * 32), smem, st, a, ncb));
  } else {
    auto kern = k_agg_tc<CB, FR, TR, -1>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(AGG_WARPS * 32), smem, st, a, ncb));
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// --------------------------------------------------------------------------------------------
// wide layers (more than 64 channels): one warp per (row point, row frame) walks ALL channel blocks of the item.
// The basis of the item's (row frame, gathered entry) pairs is evaluated once, during the first channel block, and
// its mma B fragments are kept in shared memory (the first STASH_KS k-steps = 16 * STASH_KS entries of the row;
// longer rows re-evaluate the tail); the other channel blocks only gather their 64-channel slices and run the
// aggregation mma.  k_agg_tc evaluates the basis once per channel block: 2x / 4x / 8x at 128 / 256 / 512 channels.
// --------------------------------------------------------------------------------------------
constexpr int WIDE_CB = 64;
constexpr int STASH_KS = 8;
struct WideSmem {
  static constexpr int XS = WIDE_CB + 8;
  static constexpr int X_BYTES = 32 * XS * 2;
  static constexpr int RG_BYTES = 32 * 48;
  static constexpr int G_BYTES = WIDE_CB * 64;  // T staging tile [64][64 B]; the geometry rows (32 * GROW) alias it
  static constexpr int RS_BYTES = 64;
  static constexpr int ST_BYTES = STASH_KS * 1024;  // per k-step: [2 halves][32 lanes][16 B]
  static constexpr int OFF_RG = 2 * X_BYTES;
  static constexpr int OFF_G = OFF_RG + 2 * RG_BYTES;
  static constexpr int OFF_RS = OFF_G + G_BYTES;
  static constexpr int OFF_ST = OFF_RS + 2 * RS_BYTES;
  static constexpr int WARP_BYTES = OFF_ST + ST_BYTES;
};

// one k-step (16 entries) of a 64-channel block for one row frame.  MODE 0: evaluate the basis; 1: evaluate and keep
// the fragments at st_s; 2: take them from st_s.
template <int NG, int MODE, int ACT>
__device__ __forceinline__ void wide_kstep(float (&acc)[WIDE_CB / 16][4][4], const uint32_t (&aw)[2][4], uint32_t gs_s,
                                           uint32_t st_s, const __nv_bfloat16* Xs, int ks, int lane, int act) {
  constexpr int XS = WideSmem::XS, MT = WIDE_CB / 16;
  const int g = lane >> 2, t = lane & 3, mid = lane >> 3, mr = lane & 7;
  uint32_t hb[4][2];
  if (MODE == 2) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(hb[0][0]), "=r"(hb[1][0]), "=r"(hb[2][0]), "=r"(hb[3][0]) : "r"(st_s + lane * 16) : "memory");
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(hb[0][1]), "=r"(hb[1][1]), "=r"(hb[2][1]), "=r"(hb[3][1]) : "r"(st_s + 512 + lane * 16) : "memory");
  } else {
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
      if (hq < NG) {
        float d[2][4];
        basis_pre(aw, gs_s + (ks * 16 + hq * 8) * GROW + geom_lane_off(g, t), d);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          hb[2 * m][hq] = act_pair<ACT>(d[m][0], d[m][1], act);
          hb[2 * m + 1][hq] = act_pair<ACT>(d[m][2], d[m][3], act);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) hb[j][hq] = 0u;
      }
    }
    if (MODE == 1) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_s + lane * 16), "r"(hb[0][0]), "r"(hb[1][0]),
                   "r"(hb[2][0]), "r"(hb[3][0]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_s + 512 + lane * 16), "r"(hb[0][1]), "r"(hb[1][1]),
                   "r"(hb[2][1]), "r"(hb[3][1]) : "memory");
    }
  }
  uint32_t af[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
    ldmatrix_x4_trans(af[m][0], af[m][1], af[m][2], af[m][3],
                      smem_u32(Xs + (ks * 16 + (mid >> 1) * 8 + mr) * XS + m * 16 + (mid & 1) * 8));
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int m = 0; m < MT; ++m) mma_bf16(acc[m][j], af[m], hb[j][0], hb[j][1]);
}

template <bool TR, int ACT>
__global__ void __launch_bounds__(AGG_WARPS * 32, 2) k_agg_wide(const TcAggArgs a, const int ncb) {
  using SM = WideSmem;
  constexpr int CB = WIDE_CB, MT = CB / 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + wib * SM::WARP_BYTES;
  const uint32_t wbase_s = smem_u32(wbase);
  unsigned char* Gs = wbase + SM::OFF_G;
  const uint32_t gs_s = wbase_s + SM::OFF_G, st_s = wbase_s + SM::OFF_ST;
  const int g = lane >> 2, t = lane & 3;
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int total = (int)a.n_rows * a.f_row;
  int item0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  pdl_wait();
  pdl_trigger();
  if (item0 >= total) return;
  uint32_t aw[2][4];
  load_w9_frags(a.w9, a.bias, g, t, act_pre_scale(a.act), aw);

  auto issue = [&](int item, int cb, int lo, int nt, int idx, int base, int buf, int rb, bool with_row) {
    const int gidx = chunk_gidx(a.nbr, lo, nt, idx, a.f_g, base, lane);
    gather_chunk_async<CB>(a.rec_g, a.feat, a.cs, cb * CB, gidx, lane, wbase_s + SM::OFF_RG + buf * SM::RG_BYTES,
                           wbase_s + buf * SM::X_BYTES);
    if (with_row && lane < 3)
      cp_async16_ca(wbase_s + SM::OFF_RS + rb * SM::RS_BYTES + lane * 16,
                    reinterpret_cast<const float4*>(a.rec_row) + (int64_t)item * 3 + lane, true);
  };
  auto bounds = [&](int item, int& lo, int& ne) {
    const int rp = item / a.f_row;
    lo = rp > 0 ? __ldg(a.row_ends + rp - 1) : 0;
    ne = __ldg(a.row_ends + rp) - lo;
  };
  int lo0, ne0;
  bounds(item0, lo0, ne0);
  int idx0 = lane < ne0 ? __ldg(a.nbr + lo0 + lane) : 0;
  int item1 = item0 + nwarps, lo1 = 0, ne1 = 0, idx1 = 0;
  if (item1 < total) {
    bounds(item1, lo1, ne1);
    idx1 = lane < ne1 ? __ldg(a.nbr + lo1 + lane) : 0;
  }
  int buf = 0, rb = 0;
  issue(item0, 0, lo0, ne0 * a.f_g, idx0, 0, 0, 0, true);
  cp_async_commit();

  for (;;) {
    const int nt0 = ne0 * a.f_g;
    const int nch = max(1, (nt0 + 31) >> 5);
    for (int cb = 0; cb < ncb; ++cb) {
      float acc[MT][4][4];
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[m][j][i] = 0.0f;
      for (int ch = 0; ch < nch; ++ch) {
        const int base = ch * 32;
        // ---- the next (channel block, chunk) of this item, or the first of the next item, in flight
        if (ch + 1 < nch) {
          issue(item0, cb, lo0, nt0, idx0, base + 32, buf ^ 1, rb, false);
        } else if (cb + 1 < ncb) {
          issue(item0, cb + 1, lo0, nt0, idx0, 0, buf ^ 1, rb, false);
        } else if (item1 < total) {
          issue(item1, 0, lo1, ne1 * a.f_g, idx1, 0, buf ^ 1, rb ^ 1, true);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const bool stashed = 2 * ch + 1 < STASH_KS;          // both k-steps of this chunk fit the stash
        const bool evaluate = cb == 0 || !stashed;
        if (evaluate) {
          const float* Rs = reinterpret_cast<const float*>(wbase + SM::OFF_RS + rb * SM::RS_BYTES);
          const float* rg = reinterpret_cast<const float*>(wbase + SM::OFF_RG + buf * SM::RG_BYTES) + lane * 12;
          float dx, dy, dz, Fq[9], Frow[9], gg[9];
          unpack_neighbour<TR>(rg, Rs[0], Rs[1], Rs[2], a.norm, dx, dy, dz, Fq);
          load_row_frame(Rs, 0, Frow);
          geometry9<TR>(Frow, Fq, dx, dy, dz, gg);
          store_geometry_row(gs_s + lane * GROW, lane, gg, 1.0f);
          __syncwarp();
        }
        const __nv_bfloat16* Xs = reinterpret_cast<const __nv_bfloat16*>(wbase + buf * SM::X_BYTES);
        const int n_here = min(32, nt0 - base);
#pragma unroll 1
        for (int ks = 0; ks < 2; ++ks) {
          if (ks * 16 >= n_here) break;
          const uint32_t slot = st_s + (2 * ch + ks) * 1024;
          if (!evaluate) {
            wide_kstep<2, 2, ACT>(acc, aw, gs_s, slot, Xs, ks, lane, a.act);
          } else if (ks * 16 + 8 < n_here) {
            if (stashed) wide_kstep<2, 1, ACT>(acc, aw, gs_s, slot, Xs, ks, lane, a.act);
            else wide_kstep<2, 0, ACT>(acc, aw, gs_s, slot, Xs, ks, lane, a.act);
          } else {
            if (stashed) wide_kstep<1, 1, ACT>(acc, aw, gs_s, slot, Xs, ks, lane, a.act);
            else wide_kstep<1, 0, ACT>(acc, aw, gs_s, slot, Xs, ks, lane, a.act);
          }
        }
        __syncwarp();
        buf ^= 1;
      }
      // ---- epilogue of this channel block: [64][32] bf16 through the swizzled staging tile
      {
        uint32_t* ts = reinterpret_cast<uint32_t*>(Gs);
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r0 = m * 16 + g, r1 = r0 + 8;
            ts[r0 * 16 + ((j ^ ((r0 >> 1) & 3)) << 2) + t] = pack_bf16(acc[m][j][0], acc[m][j][1]);
            ts[r1 * 16 + ((j ^ ((r1 >> 1) & 3)) << 2) + t] = pack_bf16(acc[m][j][2], acc[m][j][3]);
          }
        __syncwarp();
        const uint4* tv = reinterpret_cast<const uint4*>(Gs);
        const int ch0 = cb * CB;
        __nv_bfloat16* o = a.out + (int64_t)item0 * (int64_t)a.c * 32 + (int64_t)ch0 * 32;
#pragma unroll
        for (int it = 0; it < CB / 8; ++it) {
          const int row = it * 8 + (lane >> 2), part = lane & 3;
          if (ch0 + row < a.c) reinterpret_cast<uint4*>(o + row * 32)[part] = tv[row * 4 + (part ^ ((row >> 1) & 3))];
        }
        __syncwarp();
      }
    }
    if (item1 >= total) break;
    item0 = item1; lo0 = lo1; ne0 = ne1; idx0 = idx1;
    rb ^= 1;
    item1 += nwarps;
    if (item1 < total) {
      bounds(item1, lo1, ne1);
      idx1 = lane < ne1 ? __ldg(a.nbr + lo1 + lane) : 0;
    }
  }
  cp_async_wait<0>();
}

template <bool TR>
static int launch_agg_wide(const TcAggArgs& a, int64_t n_g, cudaStream_t st) {
  const size_t smem = AGG_WARPS * WideSmem::WARP_BYTES;
  const int ncb = (a.c + WIDE_CB - 1) / WIDE_CB;
  const int64_t warps = a.n_rows * a.f_row;
  if (warps >= (int64_t)1 << 30 || n_g * a.f_g * a.cs >= (int64_t)1 << 31 || (a.cs & 7) || a.cs < a.c) {
    set_error("launch_agg_wide: problem too large for 32-bit row offsets");
    return SE3_EINVAL;
  }
  int64_t blocks = (warps + AGG_WARPS - 1) / AGG_WARPS;
  const int64_t cap = (int64_t)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  ProfScope prof(TR ? 1 : 0, st);
  if (a.act == 2) {
    auto kern = k_agg_wide<TR, 2>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(AGG_WARPS * 32), smem, st, a, ncb));
  } else {
    auto kern = k_agg_wide<TR, -1>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(AGG_WARPS * 32), smem, st, a, ncb));
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

template <bool TR>
static int launch_agg_tc(const TcAggArgs& a0, int f_row, int64_t n_g, cudaStream_t st) {
  if (a0.n_rows == 0) return SE3_OK;
  TcAggArgs a = a0;
  a.f_row = f_row;
  // more than 16 channels with 3 / 4 row frames: items of one / two frames over 32-channel blocks (frame groups
  // repeat only the gather; 16-channel blocks would evaluate the basis once per block)
  // more than 32 channels: one row frame per item over 64-channel blocks -- the same 64 accumulators per lane as two
  // frames x 32 channels, but half the basis evaluations (the dominant cost), since every channel block of an item
  // re-evaluates the basis of the item's frames
  static const bool no_cb64 = getenv("SE3_AGG_NO_CB64") != nullptr;  // tuning aid
  static const bool no_wide = getenv("SE3_AGG_NO_WIDE") != nullptr;  // tuning aid
  if (a.c > WIDE_CB && f_row >= 1 && f_row <= 4 && !no_wide) return launch_agg_wide<TR>(a, n_g, st);
  if (a.c > 32 && f_row >= 1 && f_row <= 4 && !no_cb64) return launch_agg_cfg<64, 1, TR>(a, n_g, st);
  switch (f_row) {
    case 1: return a.c > 16 ? launch_agg_cfg<32, 1, TR>(a, n_g, st) : launch_agg_cfg<16, 1, TR>(a, n_g, st);
    case 2: return a.c > 16 ? launch_agg_cfg<32, 2, TR>(a, n_g, st) : launch_agg_cfg<16, 2, TR>(a, n_g, st);
    case 3: return a.c > 16 ? launch_agg_cfg<32, 1, TR>(a, n_g, st) : launch_agg_cfg<16, 3, TR>(a, n_g, st);
    case 4: return a.c > 16 ? launch_agg_cfg<32, 2, TR>(a, n_g, st) : launch_agg_cfg<16, 4, TR>(a, n_g, st);
  }
  set_error("launch_agg_tc: unsupported frame count");
  return SE3_EINVAL;
}

// --------------------------------------------------------------------------------------------
// gradient of proj_axes_ / proj_biases_ by output row (tensor cores)
// --------------------------------------------------------------------------------------------
struct TcEdgeArgs {
  const int* row_ends;
  const int* col_src;
  const float* rec_out;  // [n_out * f_out, 12]
  const float* rec_in;   // [n_in * f_in, 12]
  int f_in;
  const __nv_bfloat16* x;  // [n_in * f_in, cs] bf16 rows
  int c;
  int cs;
  const float* w9;
  const float* bias;
  float norm;
  int act;
  const __nv_bfloat16* dT;  // [n_out*f_out, c, 32]
  int64_t n_out;
  float* partials;          // [n_ctas, 16, 32]
  __nv_bfloat16* dxe;       // merged pass: per-entry data-gradient contributions [f_out][E * f_in][cs] + one spare row
  uint32_t dxe_frame;       // elements between the row-frame blocks of dxe (E * f_in * cs); the spare row starts at f_out * it
};

// dH^T[k, n] += sum_c dT[c, k] x[n, c] for the first NP pairs of 8-neighbour groups (straight-line code).
//   A = dT^T: [m = k][kk = c] from dTs[c][k] (.trans); B = x^T: [kk = c][nn = n] from Xs[n][c]
// FIRST: the accumulators are written, not accumulated into, by the first channel k-step (nothing to clear beforehand)
template <int CB, int NP, bool FIRST = false>
__device__ __forceinline__ void edge_dh(float (&dH)[2][4][4], const __nv_bfloat16* dTs, const __nv_bfloat16* Xs, int lane) {
  constexpr int XS = CB + 8, TS = 32 + 8;
  const int mid = lane >> 3, mr = lane & 7;
#pragma unroll
  for (int ks = 0; ks < CB / 16; ++ks) {
    uint32_t af[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
      ldmatrix_x4_trans(af[m][0], af[m][1], af[m][2], af[m][3],
                        smem_u32(dTs + (ks * 16 + (mid >> 1) * 8 + mr) * TS + m * 16 + (mid & 1) * 8));
#pragma unroll
    for (int jp = 0; jp < NP; ++jp) {
      uint32_t b[4];  // n-tiles 2jp (b0,b1) and 2jp+1 (b0,b1)
      ldmatrix_x4(b[0], b[1], b[2], b[3], smem_u32(Xs + (jp * 16 + (mid >> 1) * 8 + mr) * XS + ks * 16 + (mid & 1) * 8));
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (FIRST && ks == 0) {
          mma_bf16_zero(dH[m][2 * jp], af[m], b[0], b[1]);
          mma_bf16_zero(dH[m][2 * jp + 1], af[m], b[2], b[3]);
        } else {
          mma_bf16(dH[m][2 * jp], af[m], b[0], b[1]);
          mma_bf16(dH[m][2 * jp + 1], af[m], b[2], b[3]);
        }
      }
    }
  }
}

// dpre = dH * act'(pre) for 2 NP groups -- the pre^T tiles from the tensor cores (f16 operands) land in the accumulator
// layout of dH: k in {g, g+8} + 16 m, n = 8 j + 2 t + {0,1} -- then accA[k, d] += dpre[k, n] G[n, d] with dpre
// repacked as A fragments and G from Gb[n][16] (.trans).  Groups beyond the valid ones carry dH = 0.
template <int NP, int ACT, bool WITH_H = false>
__device__ __forceinline__ void edge_finish(float (&dH)[2][4][4], float (&accA)[2][2][4], const uint32_t (&aw)[2][4],
                                            uint32_t gs_s, const __nv_bfloat16* Gb, int lane, int act,
                                            uint32_t (*hP)[4][2] = nullptr) {
  constexpr int GB = 16 + 8;
  const int g = lane >> 2, t = lane & 3, mid = lane >> 3, mr = lane & 7;
  uint32_t dP[2][2 * NP][2];  // dpre as bf16 pairs: [m][jj][rows g / g+8]
#pragma unroll
  for (int jj = 0; jj < 2 * NP; ++jj) {
    float d[2][4];
    basis_pre(aw, gs_s + 8 * jj * GROW + geom_lane_off(g, t), d);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      if (WITH_H) {   // basis values h^T (k = 16 m + g (+8), entries 8 jj + 2t, +1) for the data-gradient product
        dP[m][jj][0] = act_both_pair<ACT>(dH[m][jj][0], dH[m][jj][1], d[m][0], d[m][1], act, hP[m][jj][0]);
        dP[m][jj][1] = act_both_pair<ACT>(dH[m][jj][2], dH[m][jj][3], d[m][2], d[m][3], act, hP[m][jj][1]);
      } else {
        dP[m][jj][0] = act_grad_pair<ACT>(dH[m][jj][0], dH[m][jj][1], d[m][0], d[m][1], act);
        dP[m][jj][1] = act_grad_pair<ACT>(dH[m][jj][2], dH[m][jj][3], d[m][2], d[m][3], act);
      }
    }
  }
#pragma unroll
  for (int ks = 0; ks < NP; ++ks) {
    uint32_t gb[4];
    ldmatrix_x4_trans(gb[0], gb[1], gb[2], gb[3], smem_u32(Gb + (ks * 16 + (mid & 1) * 8 + mr) * GB + (mid >> 1) * 8));
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const uint32_t afr[4] = {dP[m][2 * ks][0], dP[m][2 * ks][1], dP[m][2 * ks + 1][0], dP[m][2 * ks + 1][1]};
      mma_bf16(accA[m][0], afr, gb[0], gb[1]);
      mma_bf16(accA[m][1], afr, gb[2], gb[3]);
    }
  }
}

// Merged backward pass: the data-gradient contribution of every entry of the chunk for ONE row frame a,
//   dXe[n, c] = sum_k h[n, k] dT_a[c, k]   (n = (edge, in-frame b) entry, dT_a already carries out_scale),
// from the basis values the basis-gradient step has just produced.  h^T sits in the accumulator layout (rows k, pairs
// along n); the product contracts over k, so it takes one trip through shared memory: Hs[k][n] (padded rows) -> A
// fragments [m = n][kk = k] by ldmatrix.trans; B = dT^T [kk = k][nn = c] straight from the staged tile dTs[c][k].
// The first k-step writes the accumulators.
template <int CB, int NP>
__device__ __forceinline__ void edge_dx(float (&dX)[2][CB / 8][4], const uint32_t (*hP)[4][2], __nv_bfloat16* Hs,
                                        const __nv_bfloat16* dTs, int lane) {
  constexpr int HS = 40, TS = 32 + 8;
  const int g = lane >> 2, t = lane & 3, mid = lane >> 3, mr = lane & 7;
  uint32_t* hw = reinterpret_cast<uint32_t*>(Hs);
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int jj = 0; jj < 2 * NP; ++jj) {
      const int k0 = 16 * m + g, n = 8 * jj + 2 * t;
      hw[(k0 * HS + n) >> 1] = hP[m][jj][0];
      hw[((k0 + 8) * HS + n) >> 1] = hP[m][jj][1];
    }
  __syncwarp();
#pragma unroll
  for (int ki = 0; ki < 2; ++ki) {
    uint32_t af[NP][4];
#pragma unroll
    for (int mi = 0; mi < NP; ++mi)
      ldmatrix_x4_trans(af[mi][0], af[mi][1], af[mi][2], af[mi][3],
                        smem_u32(Hs + (ki * 16 + (mid >> 1) * 8 + mr) * HS + mi * 16 + (mid & 1) * 8));
#pragma unroll
    for (int jp = 0; jp < CB / 16; ++jp) {
      uint32_t b[4];
      ldmatrix_x4(b[0], b[1], b[2], b[3], smem_u32(dTs + (jp * 16 + (mid >> 1) * 8 + mr) * TS + ki * 16 + (mid & 1) * 8));
#pragma unroll
      for (int mi = 0; mi < NP; ++mi) {
        if (ki == 0) {
          mma_bf16_zero(dX[mi][2 * jp], af[mi], b[0], b[1]);
          mma_bf16_zero(dX[mi][2 * jp + 1], af[mi], b[2], b[3]);
        } else {
          mma_bf16(dX[mi][2 * jp], af[mi], b[0], b[1]);
          mma_bf16(dX[mi][2 * jp + 1], af[mi], b[2], b[3]);
        }
      }
    }
  }
}

template <int CB>
struct EdgeSmem {
  static constexpr int XS = CB + 8;   // Xs row (bf16)
  static constexpr int TS = 32 + 8;   // dTs row (bf16): [c][k]
  static constexpr int GB = 16 + 8;   // Gb row (bf16): [n][16]
  static constexpr int X_BYTES = 32 * XS * 2;
  static constexpr int RG_BYTES = 32 * 48;
  static constexpr int T_BYTES = CB * TS * 2;
  static constexpr int G_BYTES = 32 * GROW;
  static constexpr int GB_BYTES = 32 * GB * 2;
  static constexpr int RS_BYTES = 64;
  static constexpr int OFF_RG = 2 * X_BYTES;
  static constexpr int OFF_T = OFF_RG + 2 * RG_BYTES;
  static constexpr int OFF_G = OFF_T + 2 * T_BYTES;
  static constexpr int OFF_GB = OFF_G + G_BYTES;
  static constexpr int OFF_RS = OFF_GB + GB_BYTES;
  static constexpr int WARP_BYTES = OFF_RS + 2 * RS_BYTES;
};

#ifndef SE3_EDGE_MIN_BLOCKS
#define SE3_EDGE_MIN_BLOCKS 3
#endif

// One warp per (output point, output frame) item, strided over a persistent grid; per-CTA partial sums.
// Same cp.async software pipeline as k_agg_tc; a chunk here is (32 neighbours, one channel block) and its
// in-flight set also holds the [CB x 32] tile of dT.
template <int CB, int ACT>
__global__ void __launch_bounds__(AGG_WARPS * 32, SE3_EDGE_MIN_BLOCKS) k_edge_tc(const TcEdgeArgs a, const int f_out) {
  using SM = EdgeSmem<CB>;
  constexpr int TS = SM::TS, GB = SM::GB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + wib * SM::WARP_BYTES;
  const uint32_t wbase_s = smem_u32(wbase);
  const uint32_t gs_s = wbase_s + SM::OFF_G;                                  // geometry rows [32][GROW B]
  __nv_bfloat16* Gb = reinterpret_cast<__nv_bfloat16*>(wbase + SM::OFF_GB);   // [32][GB]
  const int g = lane >> 2, t = lane & 3;
  float accA[2][2][4];  // [k m-tile][d n-tile]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int dd = 0; dd < 2; ++dd)
#pragma unroll
      for (int i = 0; i < 4; ++i) accA[m][dd][i] = 0.0f;
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int ncb = (a.c + CB - 1) / CB;
  const int total = (int)a.n_out * f_out;
  int item0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  pdl_wait();
  pdl_trigger();
  if (item0 < total) {
    uint32_t aw[2][4];
    load_w9_frags(a.w9, a.bias, g, t, act_pre_scale(a.act), aw);
    auto issue = [&](int item, int lo, int nt, int idx, int j, int buf, int rb, bool with_row) {
      const int jb = j / ncb, cb = j - jb * ncb;
      const int gidx = chunk_gidx(a.col_src, lo, nt, idx, a.f_in, jb * 32, lane);
      gather_chunk_async<CB>(a.rec_in, a.x, a.cs, cb * CB, gidx, lane, wbase_s + SM::OFF_RG + buf * SM::RG_BYTES,
                             wbase_s + buf * SM::X_BYTES);
      // [CB x 32] tile of dT of this (item, channel block): CB rows of 64 B = 4 x 16 B
      const __nv_bfloat16* dTrow = a.dT + (int64_t)item * (int64_t)a.c * 32 + (int64_t)cb * CB * 32;
#pragma unroll
      for (int i0 = 0; i0 < CB * 4; i0 += 32) {
        const int i = i0 + lane;
        const int c = i >> 2, part = i & 3;
        const bool ok = cb * CB + c < a.c;
        cp_async16(wbase_s + SM::OFF_T + buf * SM::T_BYTES + (c * TS + part * 8) * 2, dTrow + (ok ? c * 32 + part * 8 : 0), ok);
      }
      if (with_row && lane < 3)
        cp_async16_ca(wbase_s + SM::OFF_RS + rb * SM::RS_BYTES + lane * 16,
                      reinterpret_cast<const float4*>(a.rec_out) + (int64_t)item * 3 + lane, true);
    };
    auto chunks_of = [&](int ne) { return max(1, (ne * a.f_in + 31) >> 5) * ncb; };

    int rp0 = item0 / f_out;
    int lo0 = rp0 > 0 ? __ldg(a.row_ends + rp0 - 1) : 0;
    int ne0 = __ldg(a.row_ends + rp0) - lo0;
    int idx0 = lane < ne0 ? __ldg(a.col_src + lo0 + lane) : 0;
    int item1 = item0 + nwarps, lo1 = 0, ne1 = 0, idx1 = 0;
    if (item1 < total) {
      const int rp1 = item1 / f_out;
      lo1 = rp1 > 0 ? __ldg(a.row_ends + rp1 - 1) : 0;
      ne1 = __ldg(a.row_ends + rp1) - lo1;
      idx1 = lane < ne1 ? __ldg(a.col_src + lo1 + lane) : 0;
    }
    int item2 = item1 + nwarps, e2a = 0, e2b = 0;
    if (item2 < total) {
      const int rp2 = item2 / f_out;
      e2a = rp2 > 0 ? __ldg(a.row_ends + rp2 - 1) : 0;
      e2b = __ldg(a.row_ends + rp2);
    }
    int buf = 0, rb = 0;
    issue(item0, lo0, ne0 * a.f_in, idx0, 0, 0, 0, true);
    cp_async_commit();

    for (;;) {
      const int nt0 = ne0 * a.f_in;
      const int nch = chunks_of(ne0);
      float dH[2][4][4];
      int nq = 0;
      for (int j = 0; j < nch; ++j) {
        if (j + 1 < nch) {
          issue(item0, lo0, nt0, idx0, j + 1, buf ^ 1, rb, false);
        } else if (item1 < total) {
          issue(item1, lo1, ne1 * a.f_in, idx1, 0, buf ^ 1, rb ^ 1, true);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const int jb = j / ncb, cb = j - jb * ncb;
        if (cb == 0) {
          const int base = jb * 32;
          const bool valid = base + lane < nt0;
          nq = (min(32, nt0 - base) + 7) >> 3;  // valid 8-neighbour groups (warp-uniform)
          const float* Rs = reinterpret_cast<const float*>(wbase + SM::OFF_RS + rb * SM::RS_BYTES);
          const float* rg = reinterpret_cast<const float*>(wbase + SM::OFF_RG + buf * SM::RG_BYTES) + lane * 12;
          float dx, dy, dz, Fq[9], Frow[9], gg[9];
          unpack_neighbour<false>(rg, Rs[0], Rs[1], Rs[2], a.norm, dx, dy, dz, Fq);
          load_row_frame(Rs, 0, Frow);
          geometry9<false>(Frow, Fq, dx, dy, dz, gg);
          if (!valid) {
#pragma unroll
            for (int i = 0; i < 9; ++i) gg[i] = 0.0f;
          }
          store_geometry_row(gs_s + lane * GROW, lane, gg, 1.0f);
          uint4 p0, p1;
          p0.x = pack_bf16(gg[0], gg[1]); p0.y = pack_bf16(gg[2], gg[3]);
          p0.z = pack_bf16(gg[4], gg[5]); p0.w = pack_bf16(gg[6], gg[7]);
          p1.x = pack_bf16(gg[8], valid ? 1.0f : 0.0f);  // column 9 = 1 -> bias gradient
          p1.y = 0u; p1.z = 0u; p1.w = 0u;
          uint4* gb = reinterpret_cast<uint4*>(Gb + lane * GB);
          gb[0] = p0;
          gb[1] = p1;
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
              for (int i = 0; i < 4; ++i) dH[m][jj][i] = 0.0f;
          __syncwarp();
        }
        // dH^T[k, n] += sum_c dT[c, k] x[n, c] over this channel block
        {
          const __nv_bfloat16* Xs = reinterpret_cast<const __nv_bfloat16*>(wbase + buf * SM::X_BYTES);
          const __nv_bfloat16* dTs = reinterpret_cast<const __nv_bfloat16*>(wbase + SM::OFF_T + buf * SM::T_BYTES);
          if (nq > 2)
            edge_dh<CB, 2>(dH, dTs, Xs, lane);
          else if (nq > 0)
            edge_dh<CB, 1>(dH, dTs, Xs, lane);
        }
        if (cb == ncb - 1) {
          if (nq > 2)
            edge_finish<2, ACT>(dH, accA, aw, gs_s, Gb, lane, a.act);
          else if (nq > 0)
            edge_finish<1, ACT>(dH, accA, aw, gs_s, Gb, lane, a.act);
        }
        __syncwarp();
        buf ^= 1;
      }
      // ---- rotate the row pipeline
      if (item1 >= total) break;
      item0 = item1; lo0 = lo1; ne0 = ne1; idx0 = idx1;
      rb ^= 1;
      item1 = item2;
      if (item1 < total) {
        lo1 = e2a;
        ne1 = e2b - e2a;
        idx1 = lane < ne1 ? __ldg(a.col_src + lo1 + lane) : 0;
      }
      item2 = item1 + nwarps;
      if (item2 < total) {
        const int rp2 = item2 / f_out;
        e2a = rp2 > 0 ? __ldg(a.row_ends + rp2 - 1) : 0;
        e2b = __ldg(a.row_ends + rp2);
      }
    }
    cp_async_wait<0>();
  }
  // per-CTA partial [d (16)][k (32)]: warps -> shared memory (the staging area is free now), summed in
  // warp order (deterministic)
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);  // [AGG_WARPS][512]
  {
    float* p = red + wib * 512;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int dd = 0; dd < 2; ++dd) {
        const int k = m * 16 + g, d = dd * 8 + 2 * t;
        p[d * 32 + k] = accA[m][dd][0];
        p[(d + 1) * 32 + k] = accA[m][dd][1];
        p[d * 32 + k + 8] = accA[m][dd][2];
        p[(d + 1) * 32 + k + 8] = accA[m][dd][3];
      }
  }
  __syncthreads();
  float* out = a.partials + (int64_t)blockIdx.x * 512;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    float s = red[i];
#pragma unroll
    for (int w = 1; w < AGG_WARPS; ++w) s += red[w * 512 + i];
    out[i] = s;
  }
}

// ---- row-item variant for c <= CB (one channel block): one warp per OUTPUT POINT; its FR frames share the
// gathered records / feature rows of every chunk (one gather instead of FR), the dT tiles of the row are
// fetched once per row -- issued as soon as the previous row's last dH product has consumed the buffer.
template <int CB, int FR>
struct EdgeRowSmem {
  static constexpr int XS = CB + 8, TS = 32 + 8, GB = 16 + 8;
  static constexpr int X_BYTES = 32 * XS * 2;
  static constexpr int RG_BYTES = 32 * 48;
  static constexpr int T_BYTES = CB * TS * 2;  // one frame's dT tile
  static constexpr int G_BYTES = 32 * GROW;
  static constexpr int GB_BYTES = 32 * GB * 2;
  static constexpr int RS_BYTES = ((FR * 48 + 63) / 64) * 64;
  static constexpr int OFF_RG = 2 * X_BYTES;
  static constexpr int OFF_T = OFF_RG + 2 * RG_BYTES;
  static constexpr int OFF_G = OFF_T + FR * T_BYTES;
  static constexpr int OFF_GB = OFF_G + G_BYTES;
  static constexpr int OFF_RS = OFF_GB + GB_BYTES;
  static constexpr int WARP_BYTES = OFF_RS + 2 * RS_BYTES;
};

// DX: the merged backward pass -- the kernel also writes, per (entry, row frame), the data-gradient contribution
// dXe = h . dT (edge_dx) to a.dxe; k_dx_segsum adds them over the transposed CSR.  The transposed aggregation pass, its
// [N F, Cout K] tile and the dx GEMM disappear, and the basis is evaluated twice per layer instead of three times.
template <int CB, int FR, int ACT, bool DX = false, int FG = 0>
__global__ void __launch_bounds__(AGG_WARPS * 32, CB >= 64 ? 2 : SE3_EDGE_MIN_BLOCKS) k_edge_row_tc(const TcEdgeArgs a) {
  using SM = EdgeRowSmem<CB, FR>;
  constexpr int TS = SM::TS, GB = SM::GB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int f_in = FG ? FG : a.f_in;
  unsigned char* wbase = smem_raw + wib * SM::WARP_BYTES;
  const uint32_t wbase_s = smem_u32(wbase);
  const uint32_t gs_s = wbase_s + SM::OFF_G;                                  // geometry rows [32][GROW B]
  __nv_bfloat16* Gb = reinterpret_cast<__nv_bfloat16*>(wbase + SM::OFF_GB);   // [32][GB]
  const int g = lane >> 2, t = lane & 3;
  float accA[2][2][4];  // [k m-tile][d n-tile]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int dd = 0; dd < 2; ++dd)
#pragma unroll
      for (int i = 0; i < 4; ++i) accA[m][dd][i] = 0.0f;
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int total = (int)a.n_out;
  int rp0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  pdl_wait();
  pdl_trigger();
  if (rp0 < total) {
    uint32_t aw[2][4];
    load_w9_frags(a.w9, a.bias, g, t, act_pre_scale(a.act), aw);
    auto issue_gather = [&](int rp, int lo, int nt, int idx, int base, int buf, int rb, bool with_row) {
      const int gidx = chunk_gidx(a.col_src, lo, nt, idx, f_in, base, lane);
      gather_chunk_async<CB>(a.rec_in, a.x, a.cs, 0, gidx, lane, wbase_s + SM::OFF_RG + buf * SM::RG_BYTES,
                             wbase_s + buf * SM::X_BYTES);
      if (with_row && lane < FR * 3)
        cp_async16_ca(wbase_s + SM::OFF_RS + rb * SM::RS_BYTES + lane * 16,
                      reinterpret_cast<const float4*>(a.rec_out) + (int64_t)rp * (FR * 3) + lane, true);
    };
    // [CB x 32] dT tile of (row, frame f): CB rows of 64 B = 4 x 16 B
    auto issue_tile = [&](int rp, int f) {
      const __nv_bfloat16* dTrow = a.dT + ((int64_t)rp * FR + f) * (int64_t)a.c * 32;
#pragma unroll
      for (int i0 = 0; i0 < CB * 4; i0 += 32) {
        const int i = i0 + lane;
        const int c = i >> 2, part = i & 3;
        const bool ok = c < a.c;
        cp_async16(wbase_s + SM::OFF_T + f * SM::T_BYTES + (c * TS + part * 8) * 2, dTrow + (ok ? c * 32 + part * 8 : 0), ok);
      }
    };

    int lo0 = rp0 > 0 ? __ldg(a.row_ends + rp0 - 1) : 0;
    int ne0 = __ldg(a.row_ends + rp0) - lo0;
    int idx0 = lane < ne0 ? __ldg(a.col_src + lo0 + lane) : 0;
    int rp1 = rp0 + nwarps, lo1 = 0, ne1 = 0, idx1 = 0;
    if (rp1 < total) {
      lo1 = __ldg(a.row_ends + rp1 - 1);
      ne1 = __ldg(a.row_ends + rp1) - lo1;
      idx1 = lane < ne1 ? __ldg(a.col_src + lo1 + lane) : 0;
    }
    int rp2 = rp1 + nwarps, e2a = 0, e2b = 0;
    if (rp2 < total) {
      e2a = __ldg(a.row_ends + rp2 - 1);
      e2b = __ldg(a.row_ends + rp2);
    }
    int buf = 0, rb = 0;
    issue_gather(rp0, lo0, ne0 * f_in, idx0, 0, 0, 0, true);
#pragma unroll
    for (int f = 0; f < FR; ++f) issue_tile(rp0, f);
    cp_async_commit();
    cp_async_commit();  // keeps the two-groups-per-iteration cadence (see the wait below)

    for (;;) {
      const int nt0 = ne0 * f_in;
      int base = 0;
      do {
        const bool last_chunk = base + 32 >= nt0;
        // group A of this iteration: the next chunk's gather
        if (!last_chunk) {
          issue_gather(rp0, lo0, nt0, idx0, base + 32, buf ^ 1, rb, false);
        } else if (rp1 < total) {
          issue_gather(rp1, lo1, ne1 * f_in, idx1, 0, buf ^ 1, rb ^ 1, true);
        }
        cp_async_commit();
        cp_async_wait<1>();  // everything but group A: this chunk's gather and this row's dT tiles have landed
        __syncwarp();
        const int nq = (min(32, nt0 - base) + 7) >> 3;  // valid 8-neighbour groups (warp-uniform)
        const float* Rs = reinterpret_cast<const float*>(wbase + SM::OFF_RS + rb * SM::RS_BYTES);
        const float* rg = reinterpret_cast<const float*>(wbase + SM::OFF_RG + buf * SM::RG_BYTES) + lane * 12;
        const __nv_bfloat16* Xs = reinterpret_cast<const __nv_bfloat16*>(wbase + buf * SM::X_BYTES);
        float dx, dy, dz, Fq[9];
        unpack_neighbour<false>(rg, Rs[0], Rs[1], Rs[2], a.norm, dx, dy, dz, Fq);
#pragma unroll 1
        for (int f = 0; f < FR; ++f) {
          {
            float Frow[9], gg[9];
            // padding lanes (beyond the row) need no masking: their feature rows are zero-filled, so dH and with it
            // dpre are exactly zero there, and their geometry is finite (a zero record against a finite row point)
            load_row_frame(Rs, f, Frow);
            geometry9<false>(Frow, Fq, dx, dy, dz, gg);
            store_geometry_row(gs_s + lane * GROW, lane, gg, 1.0f);
            uint4 p0, p1;
            p0.x = pack_bf16(gg[0], gg[1]); p0.y = pack_bf16(gg[2], gg[3]);
            p0.z = pack_bf16(gg[4], gg[5]); p0.w = pack_bf16(gg[6], gg[7]);
            p1.x = pack_bf16(gg[8], 1.0f);  // column 9 = 1 -> bias gradient
            p1.y = 0u; p1.z = 0u; p1.w = 0u;
            uint4* gb = reinterpret_cast<uint4*>(Gb + lane * GB);
            gb[0] = p0;
            gb[1] = p1;
          }
          float dH[2][4][4];   // written by the first k-step of edge_dh (the tiles of 8-neighbour groups >= nq stay unused)
          __syncwarp();
          const __nv_bfloat16* dTs = reinterpret_cast<const __nv_bfloat16*>(wbase + SM::OFF_T + f * SM::T_BYTES);
          if (nq > 2)
            edge_dh<CB, 2, true>(dH, dTs, Xs, lane);
          else if (nq > 0)
            edge_dh<CB, 1, true>(dH, dTs, Xs, lane);
          if (!DX && last_chunk && rp1 < total) {
            __syncwarp();            // every lane is done reading this frame's tile
            issue_tile(rp1, f);      // group B: the next row's tile streams in behind the rest of this chunk
          }
          uint32_t hP[DX ? 2 : 1][4][2];
          if (nq > 2)
            edge_finish<2, ACT, DX>(dH, accA, aw, gs_s, Gb, lane, a.act, hP);
          else if (nq > 0)
            edge_finish<1, ACT, DX>(dH, accA, aw, gs_s, Gb, lane, a.act, hP);
          if (DX) {
            __syncwarp();            // the geometry tiles (Gs | Gb) are free: they become the h^T tile
            float dX[2][CB / 8][4];
            __nv_bfloat16* Hs = reinterpret_cast<__nv_bfloat16*>(wbase + SM::OFF_G);
            if (nq > 2)
              edge_dx<CB, 2>(dX, hP, Hs, dTs, lane);
            else if (nq > 0)
              edge_dx<CB, 1>(dX, hP, Hs, dTs, lane);
            // rows of dXe are (row frame, entry) ordered: the entries of this chunk are consecutive rows.  Lanes whose
            // entry lies beyond the row write to the spare row instead of branching around the stores.
            const int nmi = nq > 2 ? 2 : (nq > 0 ? 1 : 0);
            const uint32_t spare = (uint32_t)FR * a.dxe_frame;
            const uint32_t row0 = (uint32_t)f * a.dxe_frame + (uint32_t)(lo0 * f_in + base + g) * (uint32_t)a.cs + 2 * t;
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              if (mi < nmi) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  const int nl = 16 * mi + 8 * hf;               // entry of the chunk owned by lanes g = 0
                  const uint32_t off = base + nl + g < nt0 ? row0 + (uint32_t)nl * (uint32_t)a.cs : spare + 2 * t;
                  uint32_t* o = reinterpret_cast<uint32_t*>(a.dxe + off);
#pragma unroll
                  for (int jn = 0; jn < CB / 8; ++jn)
                    if (8 * jn < a.cs) o[4 * jn] = pack_bf16(dX[mi][jn][2 * hf], dX[mi][jn][2 * hf + 1]);
                }
              }
            }
            if (last_chunk && rp1 < total) {
              __syncwarp();          // every lane is done reading this frame's tile
              issue_tile(rp1, f);
            }
          }
          __syncwarp();
        }
        cp_async_commit();  // group B (possibly empty)
        buf ^= 1;
        base += 32;
      } while (base < nt0);
      // ---- rotate the row pipeline
      if (rp1 >= total) break;
      rp0 = rp1; lo0 = lo1; ne0 = ne1; idx0 = idx1;
      rb ^= 1;
      rp1 = rp2;
      if (rp1 < total) {
        lo1 = e2a;
        ne1 = e2b - e2a;
        idx1 = lane < ne1 ? __ldg(a.col_src + lo1 + lane) : 0;
      }
      rp2 = rp1 + nwarps;
      if (rp2 < total) {
        e2a = __ldg(a.row_ends + rp2 - 1);
        e2b = __ldg(a.row_ends + rp2);
      }
    }
    cp_async_wait<0>();
  }
  // per-CTA partial [d (16)][k (32)]: warps -> shared memory (the staging area is free now), summed in
  // warp order (deterministic)
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);  // [AGG_WARPS][512]
  {
    float* p = red + wib * 512;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int dd = 0; dd < 2; ++dd) {
        const int k = m * 16 + g, d = dd * 8 + 2 * t;
        p[d * 32 + k] = accA[m][dd][0];
        p[(d + 1) * 32 + k] = accA[m][dd][1];
        p[d * 32 + k + 8] = accA[m][dd][2];
        p[(d + 1) * 32 + k + 8] = accA[m][dd][3];
      }
  }
  __syncthreads();
  float* out = a.partials + (int64_t)blockIdx.x * 512;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    float s = red[i];
#pragma unroll
    for (int w = 1; w < AGG_WARPS; ++w) s += red[w * 512 + i];
    out[i] = s;
  }
}

template <int CB, int FR>
static int launch_edge_row_cfg(const TcEdgeArgs& a, int n_warps, cudaStream_t st) {
  static_assert(AGG_WARPS * EdgeRowSmem<CB, FR>::WARP_BYTES >= AGG_WARPS * 512 * 4, "partials must fit the staging area");
  static_assert(EdgeRowSmem<CB, FR>::G_BYTES + EdgeRowSmem<CB, FR>::GB_BYTES >= 32 * 40 * 2, "the h^T tile aliases Gs | Gb");
  const size_t smem = AGG_WARPS * EdgeRowSmem<CB, FR>::WARP_BYTES;
  ProfScope prof(2, st);
  if (a.dxe) {
    if (CB > 32) {
      set_error("launch_edge_row_cfg: the merged pass covers at most 32 channels");
      return SE3_EINVAL;
    }
    if (a.act == 2) {
      auto kern = k_edge_row_tc<(CB > 32 ? 32 : CB), FR, 2, true>;
      SE3_SMEM_ONCE(kern, smem);
      SE3_CUDA(launch_pdl(kern, dim3((unsigned)(n_warps / AGG_WARPS)), dim3(AGG_WARPS * 32), smem, st, a));
    } else {
      auto kern = k_edge_row_tc<(CB > 32 ? 32 : CB), FR, -1, true>;
      SE3_SMEM_ONCE(kern, smem);
      SE3_CUDA(launch_pdl(kern, dim3((unsigned)(n_warps / AGG_WARPS)), dim3(AGG_WARPS * 32), smem, st, a));
    }
  } else if (a.act == 2 && a.f_in == 2) {
    auto kern = k_edge_row_tc<CB, FR, 2, false, 2>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)(n_warps / AGG_WARPS)), dim3(AGG_WARPS * 32), smem, st, a));
  } else if (a.act == 2) {
    auto kern = k_edge_row_tc<CB, FR, 2>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)(n_warps / AGG_WARPS)), dim3(AGG_WARPS * 32), smem, st, a));
  } else {
    auto kern = k_edge_row_tc<CB, FR, -1>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)(n_warps / AGG_WARPS)), dim3(AGG_WARPS * 32), smem, st, a));
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// Ordered reduction of the per-CTA partials: block b owns outputs 32b..32b+31; warp w sums partials w, w+32, ...
// (lane = output, 128-byte coalesced rows), then the 32 warp sums are added in warp order.
__global__ void __launch_bounds__(1024) k_edge_tc_reduce(const float* __restrict__ partials, int n_partials,
                                                         float* __restrict__ d_axes, float* __restrict__ d_bias) {
  __shared__ float sm[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int out = blockIdx.x * 32 + lane;
  pdl_wait();
  float s = 0.0f;
  for (int p = w; p < n_partials; p += 32) s += partials[(int64_t)p * 512 + out];
  sm[w][lane] = s;
  __syncthreads();
  if (w == 0) {
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) tot += sm[i][lane];
    if (out < 288) {
      if (d_axes) d_axes[out] = tot;
    } else if (out < 320 && d_bias) {
      d_bias[out - 288] = tot;
    }
  }
}

// Merged backward pass, second half: dx[(j f_in + b), c] = sum over the transposed row of input point j (ascending edge id --
// deterministic) and over the row frames a of dXe[edge, a, b, c].  One warp per input point; a lane owns the bf16 pairs
// lane, lane + 32 of the [f_in, cs] block, so every edge is one (or two) fully coalesced 128-byte reads per row frame.
template <int MAXW>
__global__ void __launch_bounds__(256) k_dx_segsum(const int* __restrict__ t_row_ends, const int* __restrict__ t_edge,
                                                   const uint32_t* __restrict__ dxe, int64_t frame_words, int f_out, int f_in,
                                                   int cs, int c, int64_t n_in, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int cs2 = cs >> 1, wpa = f_in * cs2;      // words (bf16 pairs) per row frame of an edge
  pdl_wait();
  for (int64_t j = gw; j < n_in; j += nw) {
    const int lo = j > 0 ? __ldg(t_row_ends + j - 1) : 0, hi = __ldg(t_row_ends + j);
    float2 acc[MAXW];
#pragma unroll
    for (int w = 0; w < MAXW; ++w) acc[w] = make_float2(0.f, 0.f);
    for (int t0 = lo; t0 < hi; t0 += 32) {
      const int mine = t0 + lane < hi ? __ldg(t_edge + t0 + lane) : 0;
      const int cnt = min(32, hi - t0);
#pragma unroll 4
      for (int i = 0; i < cnt; ++i) {
        const uint32_t* row = dxe + (int64_t)__shfl_sync(0xffffffffu, mine, i) * wpa;
        for (int a = 0; a < f_out; ++a) {
#pragma unroll
          for (int w = 0; w < MAXW; ++w) {
            const int wi = lane + 32 * w;
            if (wi < wpa) {
              const uint32_t v = __ldcs(row + a * frame_words + wi);
              acc[w].x += __uint_as_float(v << 16);
              acc[w].y += __uint_as_float(v & 0xffff0000u);
            }
          }
        }
      }
    }
#pragma unroll
    for (int w = 0; w < MAXW; ++w) {
      const int wi = lane + 32 * w;
      if (wi < wpa) {
        const int b = wi / cs2, cc = 2 * (wi - b * cs2);
        float* o = dx + (j * f_in + b) * (int64_t)c + cc;
        if (cc < c) o[0] = acc[w].x;
        if (cc + 1 < c) o[1] = acc[w].y;
      }
    }
  }
}

static int launch_dx_segsum(const int* t_row_ends, const int* t_edge, const __nv_bfloat16* dxe, int64_t frame_elems, int f_out,
                            int f_in, int cs, int c, int64_t n_in, float* dx, cudaStream_t st) {
  if (n_in == 0) return SE3_OK;
  const int wpa = f_in * (cs >> 1);
  if (wpa > 64) {
    set_error("launch_dx_segsum: more than 64 bf16 pairs per row frame");
    return SE3_EINVAL;
  }
  int64_t blocks = (n_in + 7) / 8;
  if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
  ProfScope prof(1, st);   // the data-gradient gather of the merged pass (counted where the transposed aggregation was)
  if (wpa <= 32) {
    SE3_CUDA(launch_pdl(k_dx_segsum<1>, dim3((unsigned)blocks), dim3(256), 0, st, t_row_ends, t_edge,
                        reinterpret_cast<const uint32_t*>(dxe), frame_elems / 2, f_out, f_in, cs, c, n_in, dx));
  } else {
    SE3_CUDA(launch_pdl(k_dx_segsum<2>, dim3((unsigned)blocks), dim3(256), 0, st, t_row_ends, t_edge,
                        reinterpret_cast<const uint32_t*>(dxe), frame_elems / 2, f_out, f_in, cs, c, n_in, dx));
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// CTAs per SM of the edge-gradient kernel that launch_edge_tc picks: the 64-channel row kernel runs two (at three
// its 168-register budget spills and it measured 133 vs 105 us on the dfaust level-2 -> 1 layer)
static int edge_blocks_per_sm(int c, int f_out) { return (c > 32 && c <= 64 && f_out <= 2) ? 2 : SE3_EDGE_MIN_BLOCKS; }
static int edge_tc_warps(int64_t n_items, int c, int f_out) {
  int64_t blocks = (n_items + AGG_WARPS - 1) / AGG_WARPS;
  const int64_t cap = (int64_t)num_sms() * edge_blocks_per_sm(c, f_out);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks * AGG_WARPS;
}

template <int CB>
static int launch_edge_cfg(const TcEdgeArgs& a, int f_out, int n_warps, cudaStream_t st) {
  static_assert(AGG_WARPS * EdgeSmem<CB>::WARP_BYTES >= AGG_WARPS * 512 * 4, "partials must fit the staging area");
  const size_t smem = AGG_WARPS * EdgeSmem<CB>::WARP_BYTES;
  ProfScope prof(2, st);
  if (a.act == 2) {
    auto kern = k_edge_tc<CB, 2>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)(n_warps / AGG_WARPS)), dim3(AGG_WARPS * 32), smem, st, a, f_out));
  } else {
    auto kern = k_edge_tc<CB, -1>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)(n_warps / AGG_WARPS)), dim3(AGG_WARPS * 32), smem, st, a, f_out));
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

static int launch_edge_tc(const TcEdgeArgs& a, int f_out, int64_t n_in, int n_warps, float* dA, float* dB, cudaStream_t st) {
  if (f_out < 1 || f_out > 4) {
    set_error("launch_edge_tc: unsupported frame count");
    return SE3_EINVAL;
  }
  if (n_in * a.f_in * a.cs >= (int64_t)1 << 31 || a.n_out * f_out >= (int64_t)1 << 30 || (a.cs & 7) || a.cs < a.c) {
    set_error("launch_edge_tc: problem too large for 32-bit row offsets");
    return SE3_EINVAL;
  }
  int rc;
  if (a.c <= 64 && f_out <= 2) {
    // one channel block: row items, the frames share the gather
    if (a.c <= 16)
      rc = f_out == 1 ? launch_edge_row_cfg<16, 1>(a, n_warps, st) : launch_edge_row_cfg<16, 2>(a, n_warps, st);
    else if (a.c <= 32)
      rc = f_out == 1 ? launch_edge_row_cfg<32, 1>(a, n_warps, st) : launch_edge_row_cfg<32, 2>(a, n_warps, st);
    else
      rc = f_out == 1 ? launch_edge_row_cfg<64, 1>(a, n_warps, st) : launch_edge_row_cfg<64, 2>(a, n_warps, st);
  } else {
    rc = a.c <= 16 ? launch_edge_cfg<16>(a, f_out, n_warps, st) : launch_edge_cfg<32>(a, f_out, n_warps, st);
  }
  if (rc) return rc;
  SE3_CUDA(launch_pdl(k_edge_tc_reduce, dim3(10), dim3(1024), 0, st, (const float*)a.partials, n_warps / AGG_WARPS, dA, dB));
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// --------------------------------------------------------------------------------------------
// generic bf16 GEMM: C[M,N] = alpha * A * B, fp32 accumulate
//   A_KMAJOR: A stored [M][K] (K contiguous) else [K][M];  B_KMAJOR: B stored [N][K] else [K][N]
// --------------------------------------------------------------------------------------------
template <bool A_KMAJOR, bool B_KMAJOR, bool OUT_BF16>
__global__ void __launch_bounds__(256) k_gemm_bf16(int M, int N, int K, float alpha, const __nv_bfloat16* __restrict__ A,
                                                   int64_t lda, const __nv_bfloat16* __restrict__ B, int64_t ldb,
                                                   void* __restrict__ Cout, int64_t ldc, int kchunk,
                                                   int64_t partial_stride) {
  constexpr int BM = 128, BN = 64, BK = 32, ST = 3;
  constexpr int A_ROWS = A_KMAJOR ? BM : BK, A_LD = (A_KMAJOR ? BK : BM) + 8;
  constexpr int B_ROWS = B_KMAJOR ? BN : BK, B_LD = (B_KMAJOR ? BK : BN) + 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Bs = As + ST * A_ROWS * A_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  const int k_begin = blockIdx.z * kchunk;
  const int k_end = min(K, k_begin + kchunk);
  const int ktiles = (k_end - k_begin + BK - 1) / BK;
  const int mid = lane >> 3, mr = lane & 7, g = lane >> 2, t = lane & 3;
  pdl_wait();

  auto load_stage = [&](int stage, int kt) {
    const int k0 = k_begin + kt * BK;
    __nv_bfloat16* as = As + stage * A_ROWS * A_LD;
    __nv_bfloat16* bs = Bs + stage * B_ROWS * B_LD;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * 256;
      if (A_KMAJOR) {
        const int m = c >> 2, kc = (c & 3) * 8;
        const bool p = (bm + m < M) && (k0 + kc < k_end);
        cp_async16(smem_u32(as + m * A_LD + kc), A + (int64_t)(bm + m) * lda + k0 + kc, p);
      } else {
        const int k = c >> 4, mc = (c & 15) * 8;
        const bool p = (k0 + k < k_end) && (bm + mc < M);
        cp_async16(smem_u32(as + k * A_LD + mc), A + (int64_t)(k0 + k) * lda + bm + mc, p);
      }
    }
    {
      const int c = tid;
      if (B_KMAJOR) {
        const int n = c >> 2, kc = (c & 3) * 8;
        const bool p = (bn + n < N) && (k0 + kc < k_end);
        cp_async16(smem_u32(bs + n * B_LD + kc), B + (int64_t)(bn + n) * ldb + k0 + kc, p);
      } else {
        const int k = c >> 3, nc = (c & 7) * 8;
        const bool p = (k0 + k < k_end) && (bn + nc < N);
        cp_async16(smem_u32(bs + k * B_LD + nc), B + (int64_t)(k0 + k) * ldb + bn + nc, p);
      }
    }
  };

  float acc[2][4][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][j][i] = 0.0f;

#pragma unroll
  for (int s = 0; s < ST - 1; ++s) {
    if (s < ktiles) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<ST - 2>();
    __syncthreads();
    if (kt + ST - 1 < ktiles) load_stage((kt + ST - 1) % ST, kt + ST - 1);
    cp_async_commit();
    const __nv_bfloat16* as = As + (kt % ST) * A_ROWS * A_LD;
    const __nv_bfloat16* bs = Bs + (kt % ST) * B_ROWS * B_LD;
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t af[2][4], bf[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int m0 = wm * 32 + m * 16, k0 = ks * 16;
        if (A_KMAJOR)
          ldmatrix_x4(af[m][0], af[m][1], af[m][2], af[m][3], smem_u32(as + (m0 + (mid & 1) * 8 + mr) * A_LD + k0 + (mid >> 1) * 8));
        else
          ldmatrix_x4_trans(af[m][0], af[m][1], af[m][2], af[m][3], smem_u32(as + (k0 + (mid >> 1) * 8 + mr) * A_LD + m0 + (mid & 1) * 8));
      }
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        const int n0 = wn * 32 + jp * 16, k0 = ks * 16;
        if (B_KMAJOR)
          ldmatrix_x4(bf[jp][0], bf[jp][1], bf[jp][2], bf[jp][3], smem_u32(bs + (n0 + (mid >> 1) * 8 + mr) * B_LD + k0 + (mid & 1) * 8));
        else
          ldmatrix_x4_trans(bf[jp][0], bf[jp][1], bf[jp][2], bf[jp][3], smem_u32(bs + (k0 + (mid & 1) * 8 + mr) * B_LD + n0 + (mid >> 1) * 8));
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
          mma_bf16(acc[m][2 * jp], af[m], bf[jp][0], bf[jp][1]);
          mma_bf16(acc[m][2 * jp + 1], af[m], bf[jp][2], bf[jp][3]);
        }
    }
  }
  cp_async_wait<0>();
  const float sc = partial_stride ? 1.0f : alpha;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = bm + wm * 32 + m * 16 + g + half * 8;
        const int col = bn + wn * 32 + j * 8 + 2 * t;
        if (row >= M) continue;
        const float v0 = sc * acc[m][j][half * 2], v1 = sc * acc[m][j][half * 2 + 1];
        if (OUT_BF16) {
          __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(Cout) + (int64_t)row * ldc;
          if (col + 1 < N) *reinterpret_cast<uint32_t*>(C + col) = pack_bf16(v0, v1);
          else if (col < N) C[col] = __float2bfloat16(v0);
        } else {
          float* C = reinterpret_cast<float*>(Cout) + (int64_t)blockIdx.z * partial_stride + (int64_t)row * ldc;
          if (col < N) C[col] = v0;
          if (col + 1 < N) C[col + 1] = v1;
        }
      }
}

template <bool AK, bool BK_, bool OB>
static int launch_gemm_cfg(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda,
                           const __nv_bfloat16* b, int64_t ldb, void* c, int64_t ldc, int splits, float* partials,
                           cudaStream_t st) {
  constexpr int BM = 128, BN = 64, BKK = 32, ST = 3;
  constexpr int A_ROWS = AK ? BM : BKK, A_LD = (AK ? BKK : BM) + 8;
  constexpr int B_ROWS = BK_ ? BN : BKK, B_LD = (BK_ ? BKK : BN) + 8;
  const size_t smem = (size_t)ST * (A_ROWS * A_LD + B_ROWS * B_LD) * 2;
  if (splits < 1 || partials == nullptr || OB) splits = 1;
  int kchunk = (int)((k + splits - 1) / splits);
  kchunk = (kchunk + BKK - 1) / BKK * BKK;
  splits = (int)((k + kchunk - 1) / kchunk);
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)splits);
  auto kern = k_gemm_bf16<AK, BK_, OB>;
  SE3_SMEM_ONCE(kern, smem);
  if (splits > 1) {
    SE3_CUDA(launch_pdl(kern, grid, dim3(256), smem, st, (int)m, (int)n, (int)k, alpha, a, lda, b, ldb, (void*)partials, n, kchunk, m * n));
    SE3_LAUNCH_CHECK();
    splitk_reduce_launch(partials, splits, m * n, alpha, reinterpret_cast<float*>(c), st);
    SE3_LAUNCH_CHECK();
  } else {
    SE3_CUDA(launch_pdl(kern, grid, dim3(256), smem, st, (int)m, (int)n, (int)k, alpha, a, lda, b, ldb, c, ldc, kchunk, (int64_t)0));
    SE3_LAUNCH_CHECK();
  }
  return SE3_OK;
}

// C = alpha * A[M,K] . B[N,K]^T (both K-major): tcgen05/TMEM kernel when the shape allows, else mma.sync.
// SE3_GEMM_IMPL=mma forces the mma.sync kernel (A/B testing); impl: 0 auto, 1 mma.sync, 2 tcgen05.
static int gemm_impl_env() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SE3_GEMM_IMPL");
    v = (e && e[0] == 'm') ? 1 : 0;
  }
  return v;
}

// bytes of split-K partials gemm_tn may use for an fp32-output product of this shape
size_t gemm_tn_partial_bytes(int64_t m, int64_t n, int64_t k) {
  const int s = tcgen05_gemm_splits(m, n, k);
  return s > 1 ? align_up((size_t)s * m * n * 4) : 0;
}

int gemm_tn(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
            int64_t ldb, void* c, int64_t ldc, bool out_bf16, int impl, cudaStream_t st, float* partials = nullptr) {
  if (impl == 0) impl = gemm_impl_env();
  const bool can = tcgen05_gemm_supported(m, n, k, lda, ldb);
  if (impl == 2 && !can) {
    set_error("gemm_tn: shape not supported by the tcgen05 kernel");
    return SE3_EINVAL;
  }
  if (impl != 1 && can) {
    const int splits = (partials && !out_bf16) ? tcgen05_gemm_splits(m, n, k) : 1;
    // impl 3 / auto: the persistent TMA-fed kernel (proj_tma.cu); impl 2 keeps the cp.async kernel (A/B runs)
    if (impl != 2 && tma_gemm_supported(m, n, k, lda, ldb, ldc, out_bf16, a, b, c))
      return launch_gemm_tma(m, n, k, alpha, a, lda, b, ldb, c, ldc, out_bf16, splits, partials, st);
    return launch_gemm_tcgen05(m, n, k, alpha, a, lda, b, ldb, c, ldc, out_bf16, splits, partials, st);
  }
  if (out_bf16) return launch_gemm_cfg<true, true, true>(m, n, k, alpha, a, lda, b, ldb, c, ldc, 1, nullptr, st);
  return launch_gemm_cfg<true, true, false>(m, n, k, alpha, a, lda, b, ldb, c, ldc, 1, nullptr, st);
}

// --------------------------------------------------------------------------------------------
// conversions
// --------------------------------------------------------------------------------------------
__global__ void k_f32_to_bf16(const float* __restrict__ in, int64_t n, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  const int64_t n8 = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
      uint4 p;
      p.x = pack_bf16(a.x, a.y); p.y = pack_bf16(a.z, a.w); p.z = pack_bf16(b.x, b.y); p.w = pack_bf16(b.z, b.w);
      reinterpret_cast<uint4*>(out)[i] = p;
    }
    for (int64_t i = (n8 << 3) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
      out[i] = __float2bfloat16(in[i]);
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __float2bfloat16(in[i]);
  }
}
// Forward preparation in ONE launch: blocks [0, w_blocks) write the three bf16 operand layouts of conv_weights_
// (one (input channel, 256-column chunk) slab per block, staged in shared memory so that all three stores are
// coalesced), the remaining blocks convert the feature rows x -> bf16 [rows][cs] (zero padded to cs).
__global__ void __launch_bounds__(256) k_fwd_prep(const float* __restrict__ w, int c_in, int k, int c_out, int o_chunks,
                                                  __nv_bfloat16* __restrict__ wt, __nv_bfloat16* __restrict__ wb,
                                                  __nv_bfloat16* __restrict__ wp, int w_blocks,
                                                  const float* __restrict__ x, int64_t rows, int c, int cs,
                                                  __nv_bfloat16* __restrict__ xb) {
  extern __shared__ float slab[];  // [k][oc + 1]
  const int tid = threadIdx.x;
  pdl_wait();
  if ((int)blockIdx.x < w_blocks) {
    const int ci = blockIdx.x / o_chunks, o0 = (blockIdx.x - ci * o_chunks) * 256;
    const int oc = min(256, c_out - o0), ld = oc + 1;
    const int64_t ck = (int64_t)c_in * k;
    const float* src = w + (int64_t)ci * k * c_out;
    for (int i = tid; i < k * oc; i += 256) {
      const int kk = i / oc, o = i - kk * oc;
      const float v = src[(int64_t)kk * c_out + o0 + o];
      slab[kk * ld + o] = v;
      wb[((int64_t)ci * k + kk) * c_out + o0 + o] = __float2bfloat16(v);
    }
    __syncthreads();
    for (int i = tid; i < k * oc; i += 256) {
      const int o = i / k, kk = i - o * k;
      const __nv_bfloat16 v = __float2bfloat16(slab[kk * ld + o]);
      wt[(int64_t)(o0 + o) * ck + (int64_t)ci * k + kk] = v;
      wp[((int64_t)ci * c_out + o0 + o) * k + kk] = v;
    }
    return;
  }
  const int64_t xblocks = (int64_t)gridDim.x - w_blocks;
  const int64_t stride = xblocks * 256;
  const int64_t t0 = ((int64_t)blockIdx.x - w_blocks) * 256 + tid;
  if (cs == c && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0) {
    const int64_t n8 = (rows * cs) >> 3;  // cs % 8 == 0
    for (int64_t i = t0; i < n8; i += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
      uint4 p;
      p.x = pack_bf16(a.x, a.y); p.y = pack_bf16(a.z, a.w); p.z = pack_bf16(b.x, b.y); p.w = pack_bf16(b.z, b.w);
      reinterpret_cast<uint4*>(xb)[i] = p;
    }
  } else {
    const int64_t total = rows * cs;
    for (int64_t i = t0; i < total; i += stride) {
      const int64_t r = i / cs;
      const int ch = (int)(i - r * cs);
      xb[i] = __float2bfloat16(ch < c ? x[r * c + ch] : 0.0f);
    }
  }
}

// dW[c][k][o] = dWkc[(k * cp + c)][o]: the weight gradient of a fused forward comes out in the (k, padded c) order of
// the saved tile
__global__ void k_dw_from_kc(const float* __restrict__ src, int c_in, int c_out, int cp, float* __restrict__ dw) {
  pdl_wait();
  const int64_t total = (int64_t)c_in * 32 * c_out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % c_out);
    const int k = (int)((i / c_out) & 31);
    const int c = (int)(i / ((int64_t)32 * c_out));
    dw[i] = src[((int64_t)k * cp + c) * c_out + o];
  }
}

static inline int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static int splits_for_tc(int64_t m, int64_t n, int64_t k) {
  const int64_t tiles = ((m + 127) / 128) * ((n + 63) / 64);
  int64_t s = (2ll * num_sms() + tiles - 1) / tiles;
  if (s > 64) s = 64;
  if (s * 256 > k) s = k / 256;
  if (s < 1) s = 1;
  return (int)s;
}

// --------------------------------------------------------------------------------------------
// orchestration
// --------------------------------------------------------------------------------------------
static int check_tc(const se3_conv_desc* d) {
  if (!d->rec_in || !d->rec_out) {
    set_error("precision 1 needs the packed gather records rec_in / rec_out (se3_pack_records)");
    return SE3_EINVAL;
  }
  if (d->n_in * d->f_in >= ((int64_t)1 << 30) || d->n_out * d->f_out >= ((int64_t)1 << 30)) {
    set_error("precision 1: more than 2^30 (point, frame) rows");
    return SE3_EINVAL;
  }
  if (d->c_out % 8 != 0) {
    set_error("precision 1 needs c_out to be a multiple of 8 (got %d)", d->c_out);
    return SE3_EINVAL;
  }
  return SE3_OK;
}

static inline int pad8(int c) { return (c + 7) & ~7; }

// saved-for-backward buffer of precision 1: [T bf16][x bf16 rows][Wt][Wb][Wp]
struct TcSaved {
  __nv_bfloat16 *T, *xb, *Wt, *Wb, *Wp;
  size_t bytes;
};
// forward through the fused tcgen05 kernel (conv_fused.cu)?  Then the saved tile is [R][32 * CP] in (k,c) order.
// 2 = aggregation and projection in the fused kernel, 1 = fused aggregation + stand-alone GEMM, 0 = k_agg_tc + GEMM
static int fwd_fused(const se3_conv_desc* d) {
  if (fused_supported(d->c_in, d->c_out, d->f_out, d->f_in, true)) return 2;
  return fused_supported(d->c_in, d->c_out, d->f_out, d->f_in, false) ? 1 : 0;
}
static int dx_fused(const se3_conv_desc* d) {
  if (fused_supported(d->c_out, d->c_in, d->f_in, d->f_out, true)) return 2;
  return fused_supported(d->c_out, d->c_in, d->f_in, d->f_out, false) ? 1 : 0;
}
static int64_t saved_tile_cols(const se3_conv_desc* d) {
  return fwd_fused(d) ? (int64_t)32 * fused_cp(d->c_in) : (int64_t)d->c_in * d->k;
}
static int64_t u_tile_cols(const se3_conv_desc* d) {
  return dx_fused(d) ? (int64_t)32 * fused_cp(d->c_out) : (int64_t)d->c_out * d->k;
}
// merged backward pass (edge gradient + per-entry data-gradient contributions in one gather pass, then a segmented sum
// over the transposed CSR): layers of at most 32 input channels and two output frames.  OPT-IN (SE3_BWD_MERGED=1): parity
// green, but measured slower than the transposed aggregation pass it replaces (seg_head: 322 + 196 us against
// 118 + 89 + 34 us; profiles/r02_experiments.md) -- the 198 MB of per-entry contributions leave as scattered 4-byte
// stores and the next row's dT tile can no longer be prefetched behind the basis-gradient step.
static bool merged_shape(const se3_conv_desc* d) {
  static const bool off = !(getenv("SE3_BWD_MERGED") && getenv("SE3_BWD_MERGED")[0] == '1');
  const int64_t elems = d->n_edges * d->f_out * d->f_in * (int64_t)((d->c_in + 7) & ~7);   // 32-bit element offsets
  return !off && d->c_in <= 32 && d->f_out <= 2 && d->f_in <= 4 && !fused_mode() && elems < ((int64_t)1 << 31) - 64;
}
static size_t merged_dxe_bytes(const se3_conv_desc* d) {   // [f_out][E f_in][cs] bf16 + one spare row
  return merged_shape(d) ? ((size_t)d->n_edges * d->f_out * d->f_in + 1) * (size_t)((d->c_in + 7) & ~7) * 2 : 0;
}
static TcSaved tc_saved_layout(const se3_conv_desc* d, void* base) {
  const size_t wbytes = align_up((size_t)d->c_in * d->k * d->c_out * 2);
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  TcSaved s;
  s.T = reinterpret_cast<__nv_bfloat16*>(p + off); off += align_up((size_t)d->n_out * d->f_out * saved_tile_cols(d) * 2);
  s.xb = reinterpret_cast<__nv_bfloat16*>(p + off); off += align_up((size_t)d->n_in * d->f_in * pad8(d->c_in) * 2);
  if (d->weight_cache) {
    // the three operand layouts live in the caller's per-layer cache (valid across calls until the weights change)
    char* w = reinterpret_cast<char*>(d->weight_cache);
    s.Wt = reinterpret_cast<__nv_bfloat16*>(w);
    s.Wb = reinterpret_cast<__nv_bfloat16*>(w + wbytes);
    s.Wp = reinterpret_cast<__nv_bfloat16*>(w + 2 * wbytes);
  } else {
    s.Wt = reinterpret_cast<__nv_bfloat16*>(p + off); off += wbytes;
    s.Wb = reinterpret_cast<__nv_bfloat16*>(p + off); off += wbytes;
    s.Wp = reinterpret_cast<__nv_bfloat16*>(p + off); off += wbytes;
  }
  s.bytes = off + 256;
  return s;
}
size_t conv_tc_saved_bytes(const se3_conv_desc* d) { return tc_saved_layout(d, nullptr).bytes; }
size_t conv_tc_weight_cache_bytes(const se3_conv_desc* d) { return 3 * align_up((size_t)d->c_in * d->k * d->c_out * 2) + 256; }
size_t conv_tc_fwd_workspace_bytes(const se3_conv_desc* d) {
  if (fwd_fused(d) == 2) return align_up(fused_w3_bytes(d->c_in, d->c_out)) + 256;
  if (fwd_fused(d) == 1)
    return align_up((size_t)d->c_out * saved_tile_cols(d) * 2) +
           gemm_tn_partial_bytes(d->n_out * d->f_out, d->c_out, saved_tile_cols(d)) + 256;
  return gemm_tn_partial_bytes(d->n_out * d->f_out, d->c_out, (int64_t)d->c_in * d->k) + 256;
}
size_t conv_tc_bwd_workspace_bytes(const se3_conv_desc* d) {
  const int64_t R = d->n_out * d->f_out, ck = (int64_t)d->c_in * d->k, Nf = d->n_in * d->f_in;
  size_t b = 0;
  b += align_up((size_t)R * d->c_out * 2);                                        // dy bf16
  b += align_up((size_t)R * ck * 2);                                              // dT bf16
  b += dx_fused(d) == 2 ? 0 : align_up(std::max((size_t)Nf * u_tile_cols(d) * 2, merged_dxe_bytes(d)));   // U bf16 / dXe
  {
    const int64_t tc = saved_tile_cols(d);                                        // dW partials (+ the (k,c)-ordered result)
    b += align_up((size_t)splits_for_tc(tc, d->c_out, R) * tc * d->c_out * 4);
    if (fwd_fused(d)) b += align_up((size_t)tc * d->c_out * 4);
  }
  if (dx_fused(d) == 2) b += align_up(fused_w3_bytes(d->c_out, d->c_in));        // transposed projection weights
  if (dx_fused(d) == 1) b += align_up((size_t)d->c_in * u_tile_cols(d) * 2);      // ... as a plain (k,o)-ordered matrix
  b += align_up((size_t)edge_tc_warps(d->n_out * d->f_out, d->c_in, d->f_out) * 512 * 4);                       // basis-gradient partials
  b += gemm_tn_partial_bytes(Nf, d->c_in, u_tile_cols(d));                                   // dx split-K partials
  return b + 256;
}

int conv_tc_fwd(const se3_conv_desc* d, const float* x, float* y, void* saved, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  if (int rc = check_tc(d)) return rc;
  SE3_CHECK_ARG(saved, "precision 1 needs the saved buffer");
  const int64_t R = d->n_out * d->f_out, ck = (int64_t)d->c_in * d->k;
  const int fmode = fwd_fused(d);
  const int64_t tcols = saved_tile_cols(d);
  const size_t wkc_bytes = fmode == 1 ? align_up((size_t)d->c_out * tcols * 2) : 0;
  const size_t pbytes = fmode == 2 ? 0 : gemm_tn_partial_bytes(R, d->c_out, fmode ? tcols : ck);
  if ((pbytes || wkc_bytes) && (!ws || ws_bytes < pbytes + wkc_bytes)) {
    set_error("conv_tc_fwd: workspace too small");
    return SE3_EWORKSPACE;
  }
  float* partials = pbytes ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + wkc_bytes) : nullptr;
  const TcSaved sv = tc_saved_layout(d, saved);
  const int cs = pad8(d->c_in);
  const int64_t Nf = d->n_in * d->f_in;
  {
    const int o_chunks = (d->c_out + 255) / 256;
    // weight layouts: skipped when the caller's cache is up to date
    const int w_blocks = (d->weight_cache && d->weight_cache_state == 2) ? 0 : d->c_in * o_chunks;
    const int64_t work = cs == d->c_in ? Nf * cs / 8 : Nf * cs;
    const int x_blocks = Nf > 0 ? blocks_for(work) : 0;
    const size_t smem = (size_t)d->k * (std::min(256, d->c_out) + 1) * sizeof(float);
    if (w_blocks + x_blocks > 0) {
      SE3_CUDA(launch_pdl(k_fwd_prep, dim3(w_blocks + x_blocks), dim3(256), smem, st, d->conv_weights, d->c_in, d->k, d->c_out,
                          o_chunks, sv.Wt, sv.Wb, sv.Wp, w_blocks, x, Nf, d->c_in, cs, sv.xb));
      SE3_LAUNCH_CHECK();
    }
  }
  if (fmode) {
    // geometry -> basis -> aggregation (-> projection) in one tcgen05 kernel; with the projection inside, T only leaves
    // the SM as the copy dW needs
    const size_t w3b = fmode == 2 ? align_up(fused_w3_bytes(d->c_in, d->c_out)) : wkc_bytes;
    if (!ws || ws_bytes < w3b) {
      set_error("conv_tc_fwd: workspace too small");
      return SE3_EWORKSPACE;
    }
    __nv_bfloat16* w3 = reinterpret_cast<__nv_bfloat16*>(ws);
    if (int rc = launch_w3_image(d->conv_weights, d->c_in, d->c_out, false, fmode == 1, w3, st)) return rc;
    FusedArgs f;
    f.row_ends = d->row_ends; f.nbr = d->col_src; f.rec_row = d->rec_out; f.rec_g = d->rec_in; f.f_g = d->f_in;
    f.feat = sv.xb; f.cs = cs; f.c = d->c_in; f.w9 = d->proj_axes; f.bias = d->proj_biases; f.norm = d->norm_neigh_dist;
    f.act = d->act; f.out_scale = d->out_scale; f.w3img = fmode == 2 ? reinterpret_cast<const unsigned char*>(w3) : nullptr;
    f.co = d->c_out; f.out = y; f.t_save = sv.T; f.n_rows = d->n_out; f.n_edges = d->n_edges;
    if (int rc = launch_conv_fused(f, d->f_out, false, st)) return rc;
    if (fmode == 2) return SE3_OK;
    // y[r,o] = s * sum_(k,c) T[r,(k,c)] Wkc[o,(k,c)]
    return gemm_tn(R, d->c_out, tcols, d->out_scale, sv.T, tcols, w3, tcols, y, d->c_out, false, 0, st, partials);
  }
  TcAggArgs a;
  a.row_ends = d->row_ends; a.nbr = d->col_src;
  a.rec_row = d->rec_out; a.rec_g = d->rec_in; a.f_g = d->f_in;
  a.feat = sv.xb; a.c = d->c_in; a.cs = cs; a.w9 = d->proj_axes; a.bias = d->proj_biases; a.norm = d->norm_neigh_dist;
  a.act = d->act; a.out = sv.T; a.n_rows = d->n_out;
  if (int rc = launch_agg_tc<false>(a, d->f_out, d->n_in, st)) return rc;
  // y[r,o] = s * sum_(c,k) T[r,(c,k)] Wt[o,(c,k)]   (tcgen05 / TMEM)
  return gemm_tn(R, d->c_out, ck, d->out_scale, sv.T, ck, sv.Wt, ck, y, d->c_out, false, 0, st, partials);
}

// The backward of a layer is three independent chains behind the bf16 conversion of dy -- (dW), (dT -> edge gradient),
// (transposed aggregation -> dx).  They run concurrently: two chains on forked streams, joined before the call returns,
// so the caller still sees plain stream order.  Measured on the dfaust stack (profiles/r02_experiments.md): 4.77 -> 4.40 ms
// with the fork on the coarse levels only, 4.26 ms on every layer (the HBM-bound GEMMs fill the issue gaps of the
// gather kernels).  SE3_BWD_STREAMS=0 keeps everything on the caller's stream, SE3_BWD_STREAMS_LIMIT bounds the layer
// size (edges x channels / 32) that forks.
struct SideStreams {
  cudaStream_t s[2];
  cudaEvent_t fork, join[2];
  bool ok;
};
static SideStreams* side_streams() {
  static SideStreams per_dev[16];
  static bool init[16] = {};
  static const bool off = getenv("SE3_BWD_STREAMS") && getenv("SE3_BWD_STREAMS")[0] == '0';
  if (off) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStreams& ss = per_dev[dev];
  if (!init[dev]) {
    init[dev] = true;
    ss.ok = cudaStreamCreateWithFlags(&ss.s[0], cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ss.s[1], cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.join[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.join[1], cudaEventDisableTiming) == cudaSuccess;
  }
  return ss.ok ? &ss : nullptr;
}

int conv_tc_bwd(const se3_conv_desc* d, const float* x, const float* dy, const void* saved, float* dx, float* dW,
                float* dA, float* dB, void* ws, size_t ws_bytes, cudaStream_t st_main) {
  if (int rc = check_tc(d)) return rc;
  SE3_CHECK_ARG(saved && ws, "precision 1 backward needs the forward's saved buffer and a workspace");
  SE3_CHECK_ARG(!dx || (d->t_row_ends && d->t_edge && d->t_dst), "dx needs the transposed CSR");
  const int64_t R = d->n_out * d->f_out, ck = (int64_t)d->c_in * d->k, Nf = d->n_in * d->f_in;
  const int64_t ok = (int64_t)d->c_out * d->k;
  Arena ar(ws, ws_bytes);
  __nv_bfloat16* dyb = ar.take<__nv_bfloat16>(R * d->c_out);
  __nv_bfloat16* dTb = ar.take<__nv_bfloat16>(R * ck);
  const int fdx = dx_fused(d);
  const int64_t ucols = u_tile_cols(d);
  const bool merged = merged_shape(d) && dx && (dA || dB) && !fdx;
  __nv_bfloat16* U = fdx == 2 ? nullptr : ar.take<__nv_bfloat16>(std::max((size_t)(Nf * ucols), merged_dxe_bytes(d) / 2));
  const bool ffwd = fwd_fused(d) != 0;
  const int64_t tcols = saved_tile_cols(d);
  const int splits = splits_for_tc(tcols, d->c_out, R);
  float* dWp = ar.take<float>((size_t)splits * tcols * d->c_out);
  float* dWkc = ffwd ? ar.take<float>((size_t)tcols * d->c_out) : nullptr;
  __nv_bfloat16* w3t = fdx == 2 ? ar.take<__nv_bfloat16>(align_up(fused_w3_bytes(d->c_out, d->c_in)) / 2)
                                : (fdx == 1 ? ar.take<__nv_bfloat16>((size_t)d->c_in * ucols) : nullptr);
  const int n_warps = edge_tc_warps(d->n_out * d->f_out, d->c_in, d->f_out);
  float* eg = ar.take<float>((size_t)n_warps * 512);
  const size_t dx_pbytes = gemm_tn_partial_bytes(Nf, d->c_in, ucols);
  float* dx_partials = dx_pbytes ? ar.take<float>(dx_pbytes / 4) : nullptr;
  if (!ar.ok()) { set_error("conv_tc_bwd: workspace too small"); return SE3_EWORKSPACE; }
  const TcSaved sv = tc_saved_layout(d, const_cast<void*>(saved));
  const __nv_bfloat16 *T = sv.T, *xb = sv.xb, *Wb = sv.Wb, *Wpb = sv.Wp;
  cudaStream_t st = st_main;
  SE3_CUDA(launch_pdl(k_f32_to_bf16, dim3(blocks_for(R * d->c_out / 8)), dim3(256), 0, st, dy, R * d->c_out, dyb));
  SE3_LAUNCH_CHECK();
  // fork
  SideStreams* ss = nullptr;
  {
    const int chains = (dW ? 1 : 0) + ((dA || dB) ? 1 : 0) + ((dx && !merged) ? 1 : 0);
    static const int64_t limit = getenv("SE3_BWD_STREAMS_LIMIT") ? atoll(getenv("SE3_BWD_STREAMS_LIMIT")) : ((int64_t)1 << 62);
    // per-kernel timing (se3_profile_enable) wants every kernel alone on the device: no fork while it is on
    static const int64_t floor_ = getenv("SE3_BWD_STREAMS_MIN") ? atoll(getenv("SE3_BWD_STREAMS_MIN")) : 0;
    const int64_t size = d->n_edges * (int64_t)std::max(d->c_in, d->c_out) / 32;
    if (chains >= 2 && !profile_enabled() && size < limit && size >= floor_) ss = side_streams();
  }
  if (ss) {
    SE3_CUDA(cudaEventRecord(ss->fork, st_main));
    SE3_CUDA(cudaStreamWaitEvent(ss->s[0], ss->fork, 0));
    SE3_CUDA(cudaStreamWaitEvent(ss->s[1], ss->fork, 0));
  }
  st = ss ? ss->s[0] : st_main;     // chain 1: weight gradient
  if (dW) {
    // dW[(c,k), o] = s * sum_r T[r,(c,k)] dy[r,o]
    // tcgen05 (MN-major operands); SE3_DW_IMPL=m / SE3_GEMM_IMPL=mma select the mma.sync kernel
    static const bool dw_mma = getenv("SE3_DW_IMPL") && getenv("SE3_DW_IMPL")[0] == 'm';  // A/B aid
    // (fused forward: the saved rows are (k, c padded to CP) ordered -> the result is permuted afterwards)
    float* dst = ffwd ? dWkc : dW;
    if (!dw_mma && gemm_impl_env() != 1 && tma_gemm_mn_supported(tcols, d->c_out, R, tcols, d->c_out, d->c_out, T, dyb, dst)) {
      // persistent TMA-fed kernel (producer / MMA / epilogue warps, double-buffered TMEM accumulator)
      if (int rc = launch_gemm_tma_mn(tcols, d->c_out, R, d->out_scale, T, tcols, dyb, d->c_out, dst, d->c_out, splits, dWp, st))
        return rc;
    } else if (!dw_mma && gemm_impl_env() != 1 && tcgen05_gemm_mn_supported(tcols, d->c_out, R, tcols, d->c_out)) {
      if (int rc = launch_gemm_tcgen05_mn(tcols, d->c_out, R, d->out_scale, T, tcols, dyb, d->c_out, dst, d->c_out, splits, dWp, st))
        return rc;
    } else if (int rc = launch_gemm_cfg<false, false, false>(tcols, d->c_out, R, d->out_scale, T, tcols, dyb, d->c_out, dst,
                                                             d->c_out, splits, dWp, st)) {
      return rc;
    }
    if (ffwd) {
      const int64_t n = ck * d->c_out;
      SE3_CUDA(launch_pdl(k_dw_from_kc, dim3(blocks_for(n)), dim3(256), 0, st, (const float*)dWkc, d->c_in, d->c_out,
                          fused_cp(d->c_in), dW));
      SE3_LAUNCH_CHECK();
    }
  }
  st = st_main;                     // chain 2: basis gradient (stays on the caller's stream)
  if (dA || dB) {
    // dT[r,(c,k)] = s * sum_o dy[r,o] W[(c,k),o]
    if (int rc = gemm_tn(R, ck, d->c_out, d->out_scale, dyb, d->c_out, Wb, d->c_out, dTb, ck, true, 0, st)) return rc;
    TcEdgeArgs g;
    g.row_ends = d->row_ends; g.col_src = d->col_src;
    g.rec_out = d->rec_out; g.rec_in = d->rec_in; g.f_in = d->f_in;
    g.x = xb; g.c = d->c_in; g.cs = pad8(d->c_in); g.w9 = d->proj_axes; g.bias = d->proj_biases; g.norm = d->norm_neigh_dist;
    g.act = d->act; g.dT = dTb; g.n_out = d->n_out; g.partials = eg;
    g.dxe = merged ? U : nullptr;
    g.dxe_frame = merged ? (uint32_t)(d->n_edges * d->f_in * pad8(d->c_in)) : 0u;
    if (int rc = launch_edge_tc(g, d->f_out, d->n_in, n_warps, dA, dB, st)) return rc;
    if (merged)
      if (int rc = launch_dx_segsum(d->t_row_ends, d->t_edge, U, (int64_t)g.dxe_frame, d->f_out, d->f_in, pad8(d->c_in), d->c_in,
                                    d->n_in, dx, st))
        return rc;
  }
  st = ss ? ss->s[1] : st_main;     // chain 3: data gradient
  if (merged) {
    // dx came out of the basis-gradient chain
  } else if (dx && fdx) {
    // data gradient: the same fused kernel over the transposed CSR, gathered rows = dy, projection with W^T
    if (int rc = launch_w3_image(d->conv_weights, d->c_in, d->c_out, true, fdx == 1, w3t, st)) return rc;
    FusedArgs f;
    f.row_ends = d->t_row_ends; f.nbr = d->t_dst; f.rec_row = d->rec_in; f.rec_g = d->rec_out; f.f_g = d->f_out;
    f.feat = dyb; f.cs = d->c_out; f.c = d->c_out; f.w9 = d->proj_axes; f.bias = d->proj_biases; f.norm = d->norm_neigh_dist;
    f.act = d->act; f.out_scale = d->out_scale; f.w3img = fdx == 2 ? reinterpret_cast<const unsigned char*>(w3t) : nullptr;
    f.co = d->c_in; f.out = dx; f.t_save = fdx == 2 ? nullptr : U; f.n_rows = d->n_in; f.n_edges = d->n_edges;
    if (int rc = launch_conv_fused(f, d->f_in, true, st)) return rc;
    // dx[n,c] = s * sum_(k,o) U[n,(k,o)] Wp[c,(k,o)]
    if (fdx == 1)
      if (int rc = gemm_tn(Nf, d->c_in, ucols, d->out_scale, U, ucols, w3t, ucols, dx, d->c_in, false, 0, st, dx_partials)) return rc;
  } else if (dx) {
    TcAggArgs a;
    a.row_ends = d->t_row_ends; a.nbr = d->t_dst;
    a.rec_row = d->rec_in; a.rec_g = d->rec_out; a.f_g = d->f_out;
    a.feat = dyb; a.c = d->c_out; a.cs = d->c_out; a.w9 = d->proj_axes; a.bias = d->proj_biases; a.norm = d->norm_neigh_dist;
    a.act = d->act; a.out = U; a.n_rows = d->n_in;
    if (int rc = launch_agg_tc<true>(a, d->f_in, d->n_out, st)) return rc;
    // dx[n,c] = s * sum_(o,k) U[n,(o,k)] Wp[c,(o,k)]
    if (int rc = gemm_tn(Nf, d->c_in, ok, d->out_scale, U, ok, Wpb, ok, dx, d->c_in, false, 0, st, dx_partials)) return rc;
  }
  if (ss) {   // join
    SE3_CUDA(cudaEventRecord(ss->join[0], ss->s[0]));
    SE3_CUDA(cudaEventRecord(ss->join[1], ss->s[1]));
    SE3_CUDA(cudaStreamWaitEvent(st_main, ss->join[0], 0));
    SE3_CUDA(cudaStreamWaitEvent(st_main, ss->join[1], 0));
  }
  return SE3_OK;
}

}  // namespace se3

namespace se3 {
__global__ void k_pack_records(const float* __restrict__ pts, const float* __restrict__ frames, int64_t n, int f,
                               float4* __restrict__ rec) {
  const int64_t total = n * f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / f;
    const float* fr = frames + i * 9;
    rec[3 * i] = make_float4(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], fr[0]);
    rec[3 * i + 1] = make_float4(fr[1], fr[2], fr[3], fr[4]);
    rec[3 * i + 2] = make_float4(fr[5], fr[6], fr[7], fr[8]);
  }
}
}  // namespace se3

extern "C" int se3_conv_set_fused(int32_t mode) {
  const int prev = se3::fused_mode();
  se3::fused_set_mode(mode);
  return prev;
}

extern "C" int se3_pack_records(const float* pts, const float* frames, int64_t n, int32_t f, float* rec,
                                se3_stream_t stream) {
  using namespace se3;
  SE3_CHECK_ARG(n >= 0 && f >= 1 && f <= 4, "bad sizes");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(pts && frames && rec, "null pointer");
  SE3_CHECK_ARG((reinterpret_cast<uintptr_t>(rec) & 15) == 0, "rec must be 16-byte aligned");
  k_pack_records<<<blocks_for(n * f), 256, 0, as_stream(stream)>>>(pts, frames, n, f, reinterpret_cast<float4*>(rec));
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// C = alpha * A[M,K] . B[N,K]^T, bf16 operands (K-major), fp32 or bf16 output.  impl: 0 auto, 1 mma.sync, 2 tcgen05.
extern "C" int se3_gemm_bf16_tn(const void* a, const void* b, int64_t m, int64_t n, int64_t k, float alpha, void* c,
                                int32_t c_is_bf16, int32_t impl, se3_stream_t stream) {
  using namespace se3;
  SE3_CHECK_ARG(m >= 0 && n >= 1 && k >= 8 && (k % 8) == 0, "bad sizes (k must be a multiple of 8)");
  SE3_CHECK_ARG(impl >= 0 && impl <= 3, "impl must be 0 (auto), 1 (mma.sync), 2 (tcgen05, cp.async) or 3 (tcgen05, TMA)");
  if (m == 0) return SE3_OK;
  SE3_CHECK_ARG(a && b && c, "null pointer");
  return gemm_tn(m, n, k, alpha, reinterpret_cast<const __nv_bfloat16*>(a), k, reinterpret_cast<const __nv_bfloat16*>(b), k,
                 c, n, c_is_bf16 != 0, impl, as_stream(stream));
}

// C[M,N] = alpha * A^T B, A stored [K][M], B stored [K][N] (bf16), fp32 output.  impl as in the header.
extern "C" int se3_gemm_bf16_mn(const void* a, const void* b, int64_t m, int64_t n, int64_t k, float alpha, float* c,
                                int32_t impl, se3_stream_t stream) {
  using namespace se3;
  SE3_CHECK_ARG(m >= 0 && n >= 8 && (n % 8) == 0 && (m % 8) == 0 && k >= 1, "bad sizes (m and n must be multiples of 8)");
  SE3_CHECK_ARG(impl >= 0 && impl <= 3, "impl must be 0 (auto), 1 (mma.sync), 2 (tcgen05, cp.async) or 3 (tcgen05, TMA)");
  if (m == 0) return SE3_OK;
  SE3_CHECK_ARG(a && b && c, "null pointer");
  const __nv_bfloat16 *A = reinterpret_cast<const __nv_bfloat16*>(a), *B = reinterpret_cast<const __nv_bfloat16*>(b);
  cudaStream_t st = as_stream(stream);
  const bool tma_ok = tma_gemm_mn_supported(m, n, k, m, n, n, a, b, c);
  if (impl == 3 && !tma_ok) {
    set_error("se3_gemm_bf16_mn: shape / alignment not supported by the TMA kernel");
    return SE3_EINVAL;
  }
  if ((impl == 0 || impl == 3) && tma_ok) return launch_gemm_tma_mn(m, n, k, alpha, A, m, B, n, c, n, 1, nullptr, st);
  if (impl != 1 && tcgen05_gemm_mn_supported(m, n, k, m, n)) return launch_gemm_tcgen05_mn(m, n, k, alpha, A, m, B, n, c, n, 1, nullptr, st);
  return launch_gemm_cfg<false, false, false>(m, n, k, alpha, A, m, B, n, c, n, 1, nullptr, st);
}
