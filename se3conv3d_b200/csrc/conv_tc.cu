// bf16 tensor-core path (precision 1) of the fused PNEConvLayerRotEquiv.
//
//   k_agg_tc   : one warp per output row-point.  Per chunk of 32 expanded neighbours (edge x in-frame):
//                lanes gather coordinates / frames / feature rows (128-bit loads, bf16 staging in
//                shared memory), build the 9-vector g per (neighbour, out-frame), evaluate the basis
//                h = act(g.W9 + b) DIRECTLY in the mma B-fragment layout (no shared-memory round trip
//                for h) and accumulate T[c,k] += x[n,c] h[n,k] with mma.sync m16n8k16 (bf16 in, fp32
//                accumulate).  g, h and the expanded neighbour list never reach HBM.
//                TR=false: rows = output points (forward).  TR=true: rows = input points over the
//                transposed CSR with dy as the gathered feature (atomic-free data gradient).
//   k_edge_tc  : by output row, gradient of proj_axes_/proj_biases_: dH^T = dT^T X^T (mma), times
//                act'(pre) recomputed in the accumulator layout, then [k x n].[n x 10] (mma) against
//                the bf16 geometry (+ a ones column for the bias); per-warp partials, ordered reduce.
//   k_gemm_bf16: generic bf16 tensor-core GEMM (cp.async 3-stage pipeline, ldmatrix, mma.sync) for the
//                projection and its three backward products; deterministic split-K.
// The tcgen05/TMEM projection kernel lives in proj_tcgen05.cu and replaces k_gemm_bf16 for the
// K-major x K-major products when it is enabled.
#include <stdlib.h>
#include "conv_simt.cuh"
#include "tc_common.cuh"

namespace se3 {

void splitk_reduce_launch(const float* partials, int splits, int64_t mn, float alpha, float* c, cudaStream_t st);
// proj_tcgen05.cu
bool tcgen05_gemm_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb);
int launch_gemm_tcgen05(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda,
                        const __nv_bfloat16* b, int64_t ldb, void* c, int64_t ldc, bool out_bf16, cudaStream_t st);

struct TcAggArgs {
  const int* row_ends;
  const int* nbr;
  const float* pts_row;
  const float* frm_row;
  const float* pts_g;
  const float* frm_g;
  int f_g;
  const float* feat;
  int c;
  const float* w9;
  const float* bias;
  float norm;
  int act;
  __nv_bfloat16* out;
  int64_t n_rows;
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

// Stage the gathered feature rows of a 32-neighbour chunk into Xs[32][CB+8] (bf16).
template <int CB>
__device__ __forceinline__ void stage_features(const float* __restrict__ feat, int C, int c0, int fidx, int lane,
                                               __nv_bfloat16* Xs) {
  constexpr int XS = CB + 8;
  if ((C & 3) == 0) {
    constexpr int LPR = CB / 4;    // lanes per row (float4 each)
    constexpr int RPI = 32 / LPR;  // rows per iteration
#pragma unroll
    for (int it = 0; it < LPR; ++it) {
      const int row = it * RPI + lane / LPR;
      const int col4 = lane % LPR;
      const int src = __shfl_sync(0xffffffffu, fidx, row);
      const int ch = c0 + col4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src >= 0 && ch < C) v = __ldg(reinterpret_cast<const float4*>(feat + (int64_t)src * C + ch));
      uint2 p;
      p.x = pack_bf16(v.x, v.y);
      p.y = pack_bf16(v.z, v.w);
      *reinterpret_cast<uint2*>(Xs + row * XS + col4 * 4) = p;
    }
  } else {
    for (int row = 0; row < 32; ++row) {
      const int src = __shfl_sync(0xffffffffu, fidx, row);
      for (int ch = lane; ch < CB; ch += 32) {
        float v = 0.f;
        if (src >= 0 && c0 + ch < C) v = __ldg(feat + (int64_t)src * C + c0 + ch);
        Xs[row * XS + ch] = __float2bfloat16(v);
      }
    }
  }
}

constexpr int AGG_WARPS = 4;  // warps per CTA of the aggregation / edge kernels

template <int CB, int FR, bool TR, int ACT>
__global__ void __launch_bounds__(AGG_WARPS * 32, 3) k_agg_tc(const TcAggArgs a, const int ncb) {
  constexpr int XS = CB + 8;
  constexpr int MT = CB / 16;
  constexpr int WARP_BYTES = 32 * XS * 2 + FR * 32 * 12 * 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* W9s = reinterpret_cast<float*>(smem_raw);  // [32][12]: w0..w8, bias, 0, 0
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + 32 * 12 * 4 + wib * WARP_BYTES;
  __nv_bfloat16* Xs = reinterpret_cast<__nv_bfloat16*>(wbase);
  float* Gs = reinterpret_cast<float*>(wbase + 32 * XS * 2);  // [FR][32][12]
  for (int i = threadIdx.x; i < 32 * 12; i += blockDim.x) {
    const int k = i / 12, d = i % 12;
    W9s[i] = d < 9 ? a.w9[d * 32 + k] : (d == 9 ? a.bias[k] : 0.0f);
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t total = a.n_rows * ncb;
  const int mid = lane >> 3, mr = lane & 7;  // ldmatrix: matrix id, row inside the matrix
  for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += nwarps) {
    const int64_t rp = item / ncb;
    const int c0 = (int)(item - rp * ncb) * CB;
    const int lo = rp > 0 ? a.row_ends[rp - 1] : 0;
    const int hi = a.row_ends[rp];
    const int n_total = (hi - lo) * a.f_g;
    const float prx = a.pts_row[3 * rp], pry = a.pts_row[3 * rp + 1], prz = a.pts_row[3 * rp + 2];
    float acc[FR][MT][4][4];
#pragma unroll
    for (int f = 0; f < FR; ++f)
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[f][m][j][i] = 0.0f;
    for (int base = 0; base < n_total; base += 32) {
      const int n = base + lane;
      const bool valid = n < n_total;
      const int e = valid ? n / a.f_g : 0;
      const int fg = valid ? n - e * a.f_g : 0;
      const int q = __ldg(a.nbr + lo + e);
      const int fidx = valid ? q * a.f_g + fg : -1;
      const float pgx = __ldg(a.pts_g + 3 * (int64_t)q), pgy = __ldg(a.pts_g + 3 * (int64_t)q + 1),
                  pgz = __ldg(a.pts_g + 3 * (int64_t)q + 2);
      float Fq[9];
      {
        const float* fp = a.frm_g + ((int64_t)q * a.f_g + fg) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) Fq[i] = __ldg(fp + i);
      }
      const float dx = (TR ? (prx - pgx) : (pgx - prx)) * a.norm;
      const float dy = (TR ? (pry - pgy) : (pgy - pry)) * a.norm;
      const float dz = (TR ? (prz - pgz) : (pgz - prz)) * a.norm;
#pragma unroll
      for (int f = 0; f < FR; ++f) {
        float Frow[9];
        const float* fp = a.frm_row + (rp * FR + f) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) Frow[i] = __ldg(fp + i);
        float gg[9];
        geometry9<TR>(Frow, Fq, dx, dy, dz, gg);
        float4* gs = reinterpret_cast<float4*>(Gs + (f * 32 + lane) * 12);
        gs[0] = make_float4(gg[0], gg[1], gg[2], gg[3]);
        gs[1] = make_float4(gg[4], gg[5], gg[6], gg[7]);
        gs[2] = make_float4(gg[8], 0.f, 0.f, 0.f);
      }
      stage_features<CB>(a.feat, a.c, c0, fidx, lane, Xs);
      __syncwarp();
#pragma unroll 1
      for (int ks = 0; ks < 2; ++ks) {
        // A fragments (x^T): [m = channel][k = neighbour], from Xs[n][c] through ldmatrix.trans
        uint32_t af[MT][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
          ldmatrix_x4_trans(af[m][0], af[m][1], af[m][2], af[m][3],
                            smem_u32(Xs + (ks * 16 + (mid >> 1) * 8 + mr) * XS + m * 16 + (mid & 1) * 8));
#pragma unroll
        for (int f = 0; f < FR; ++f) {
          // this lane's 4 neighbours of the k-step: 2t, 2t+1, 2t+8, 2t+9
          float gv[4][9];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int nn = ks * 16 + 2 * t + (u & 1) + (u >> 1) * 8;
            const float4* gs = reinterpret_cast<const float4*>(Gs + (f * 32 + nn) * 12);
            const float4 v0 = gs[0], v1 = gs[1], v2 = gs[2];
            gv[u][0] = v0.x; gv[u][1] = v0.y; gv[u][2] = v0.z; gv[u][3] = v0.w;
            gv[u][4] = v1.x; gv[u][5] = v1.y; gv[u][6] = v1.z; gv[u][7] = v1.w;
            gv[u][8] = v2.x;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // basis column k = g + 8 j of this lane: weights from shared memory (w0..w8, bias)
            const float4* wp = reinterpret_cast<const float4*>(W9s + (g + 8 * j) * 12);
            const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
            const float wv[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x};
            float h[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float pre = w2.y;
#pragma unroll
              for (int d = 0; d < 9; ++d) pre = fmaf(gv[u][d], wv[d], pre);
              h[u] = act_rt<ACT>(pre, a.act);
            }
            const uint32_t b0 = pack_bf16(h[0], h[1]), b1 = pack_bf16(h[2], h[3]);
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_bf16(acc[f][m][j], af[m], b0, b1);
          }
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int f = 0; f < FR; ++f) {
      __nv_bfloat16* o = a.out + (rp * FR + f) * (int64_t)a.c * 32;
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + m * 16 + g;
          const int k = j * 8 + 2 * t;
          if (c < a.c) *reinterpret_cast<uint32_t*>(o + c * 32 + k) = pack_bf16(acc[f][m][j][0], acc[f][m][j][1]);
          if (c + 8 < a.c) *reinterpret_cast<uint32_t*>(o + (c + 8) * 32 + k) = pack_bf16(acc[f][m][j][2], acc[f][m][j][3]);
        }
    }
  }
}

template <int CB, int FR, bool TR>
static int launch_agg_cfg(const TcAggArgs& a, cudaStream_t st) {
  constexpr int XS = CB + 8;
  constexpr int WARP_BYTES = 32 * XS * 2 + FR * 32 * 12 * 4;
  const size_t smem = 32 * 12 * 4 + AGG_WARPS * WARP_BYTES;
  const int ncb = (a.c + CB - 1) / CB;
  const int64_t warps = a.n_rows * ncb;
  int64_t blocks = (warps + AGG_WARPS - 1) / AGG_WARPS;
  const int64_t cap = (int64_t)num_sms() * 3 * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (a.act == 2) {
    auto kern = k_agg_tc<CB, FR, TR, 2>;
    SE3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(int)blocks, AGG_WARPS * 32, smem, st>>>(a, ncb);
  } else {
    auto kern = k_agg_tc<CB, FR, TR, -1>;
    SE3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(int)blocks, AGG_WARPS * 32, smem, st>>>(a, ncb);
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

template <bool TR>
static int launch_agg_tc(const TcAggArgs& a, int f_row, cudaStream_t st) {
  if (a.n_rows == 0) return SE3_OK;
  switch (f_row) {
    case 1: return a.c > 32 ? launch_agg_cfg<64, 1, TR>(a, st) : (a.c > 16 ? launch_agg_cfg<32, 1, TR>(a, st) : launch_agg_cfg<16, 1, TR>(a, st));
    case 2: return a.c > 16 ? launch_agg_cfg<32, 2, TR>(a, st) : launch_agg_cfg<16, 2, TR>(a, st);
    case 3: return launch_agg_cfg<16, 3, TR>(a, st);
    case 4: return launch_agg_cfg<16, 4, TR>(a, st);
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
  const float* pts_out;
  const float* frm_out;
  const float* pts_in;
  const float* frm_in;
  int f_in;
  const float* x;
  int c;
  const float* w9;
  const float* bias;
  float norm;
  int act;
  const __nv_bfloat16* dT;  // [n_out*FR, c, 32]
  int64_t n_out;
  float* partials;          // [n_warps, 16, 32]
};

template <int CB, int FR, int ACT>
__global__ void __launch_bounds__(AGG_WARPS * 32, 3) k_edge_tc(const TcEdgeArgs a) {
  constexpr int XS = CB + 8;   // Xs row (bf16)
  constexpr int TS = 32 + 8;   // dTs row (bf16): [c][k]
  constexpr int GB = 16 + 8;   // Gb row (bf16): [n][16]
  constexpr int WARP_BYTES = 32 * XS * 2 + CB * TS * 2 + 32 * 12 * 4 + 32 * GB * 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* W9s = reinterpret_cast<float*>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + 32 * 12 * 4 + wib * WARP_BYTES;
  __nv_bfloat16* Xs = reinterpret_cast<__nv_bfloat16*>(wbase);
  __nv_bfloat16* dTs = reinterpret_cast<__nv_bfloat16*>(wbase + 32 * XS * 2);
  float* Gs = reinterpret_cast<float*>(wbase + 32 * XS * 2 + CB * TS * 2);
  __nv_bfloat16* Gb = reinterpret_cast<__nv_bfloat16*>(wbase + 32 * XS * 2 + CB * TS * 2 + 32 * 12 * 4);
  for (int i = threadIdx.x; i < 32 * 12; i += blockDim.x) {
    const int k = i / 12, d = i % 12;
    W9s[i] = d < 9 ? a.w9[d * 32 + k] : (d == 9 ? a.bias[k] : 0.0f);
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int mid = lane >> 3, mr = lane & 7;
  float wr[4][9], br[4];  // k = g + 8*j  (j = 2*mtile + half)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int d = 0; d < 9; ++d) wr[j][d] = W9s[(g + 8 * j) * 12 + d];
    br[j] = W9s[(g + 8 * j) * 12 + 9];
  }
  float accA[2][2][4];  // [k m-tile][d n-tile]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int dd = 0; dd < 2; ++dd)
#pragma unroll
      for (int i = 0; i < 4; ++i) accA[m][dd][i] = 0.0f;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int ncb = (a.c + CB - 1) / CB;
  for (int64_t rp = gw; rp < a.n_out; rp += nwarps) {
    const int lo = rp > 0 ? a.row_ends[rp - 1] : 0;
    const int hi = a.row_ends[rp];
    const int n_total = (hi - lo) * a.f_in;
    const float prx = a.pts_out[3 * rp], pry = a.pts_out[3 * rp + 1], prz = a.pts_out[3 * rp + 2];
    for (int base = 0; base < n_total; base += 32) {
      const int n = base + lane;
      const bool valid = n < n_total;
      const int e = valid ? n / a.f_in : 0;
      const int fg = valid ? n - e * a.f_in : 0;
      const int q = __ldg(a.col_src + lo + e);
      const int fidx = valid ? q * a.f_in + fg : -1;
      float Fq[9];
      {
        const float* fp = a.frm_in + ((int64_t)q * a.f_in + fg) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) Fq[i] = __ldg(fp + i);
      }
      const float dx = (__ldg(a.pts_in + 3 * (int64_t)q) - prx) * a.norm;
      const float dy = (__ldg(a.pts_in + 3 * (int64_t)q + 1) - pry) * a.norm;
      const float dz = (__ldg(a.pts_in + 3 * (int64_t)q + 2) - prz) * a.norm;
#pragma unroll 1
      for (int f = 0; f < FR; ++f) {
        float Frow[9];
        {
          const float* fp = a.frm_out + (rp * FR + f) * 9;
#pragma unroll
          for (int i = 0; i < 9; ++i) Frow[i] = __ldg(fp + i);
        }
        float gg[9];
        geometry9<false>(Frow, Fq, dx, dy, dz, gg);
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 9; ++i) gg[i] = 0.0f;
        }
        {
          float4* gs = reinterpret_cast<float4*>(Gs + lane * 12);
          gs[0] = make_float4(gg[0], gg[1], gg[2], gg[3]);
          gs[1] = make_float4(gg[4], gg[5], gg[6], gg[7]);
          gs[2] = make_float4(gg[8], 0.f, 0.f, 0.f);
          uint4 p0, p1;
          p0.x = pack_bf16(gg[0], gg[1]); p0.y = pack_bf16(gg[2], gg[3]);
          p0.z = pack_bf16(gg[4], gg[5]); p0.w = pack_bf16(gg[6], gg[7]);
          p1.x = pack_bf16(gg[8], valid ? 1.0f : 0.0f);  // column 9 = 1 -> bias gradient
          p1.y = 0u; p1.z = 0u; p1.w = 0u;
          uint4* gb = reinterpret_cast<uint4*>(Gb + lane * GB);
          gb[0] = p0;
          gb[1] = p1;
        }
        // dH^T[k, n] = sum_c dT[c, k] x[n, c], accumulated over channel blocks
        float dH[2][4][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) dH[m][j][i] = 0.0f;
        const __nv_bfloat16* dTrow = a.dT + (rp * FR + f) * (int64_t)a.c * 32;
        for (int cb = 0; cb < ncb; ++cb) {
          const int c0 = cb * CB;
          __syncwarp();
          stage_features<CB>(a.x, a.c, c0, fidx, lane, Xs);
          // dT rows of this channel block: CB rows of 32 bf16 (64 B) = 4 x 16 B each
          for (int i = lane; i < CB * 4; i += 32) {
            const int c = i >> 2, part = i & 3;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (c0 + c < a.c) v = __ldg(reinterpret_cast<const uint4*>(dTrow + (int64_t)(c0 + c) * 32) + part);
            *reinterpret_cast<uint4*>(dTs + c * TS + part * 8) = v;
          }
          __syncwarp();
#pragma unroll
          for (int ks = 0; ks < CB / 16; ++ks) {
            // A = dT^T: [m = k][kk = c] from dTs[c][k] (.trans); B = x^T: [kk = c][nn = n] from Xs[n][c]
            uint32_t af[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
              ldmatrix_x4_trans(af[m][0], af[m][1], af[m][2], af[m][3],
                                smem_u32(dTs + (ks * 16 + (mid >> 1) * 8 + mr) * TS + m * 16 + (mid & 1) * 8));
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
              uint32_t b[4];  // n-tiles 2jp (b0,b1) and 2jp+1 (b0,b1)
              ldmatrix_x4(b[0], b[1], b[2], b[3], smem_u32(Xs + (jp * 16 + (mid >> 1) * 8 + mr) * XS + ks * 16 + (mid & 1) * 8));
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                mma_bf16(dH[m][2 * jp], af[m], b[0], b[1]);
                mma_bf16(dH[m][2 * jp + 1], af[m], b[2], b[3]);
              }
            }
          }
        }
        // dpre = dH * act'(pre) in the accumulator layout: k in {g, g+8} + 16 m, n = 8 j + 2 t + {0,1}
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float gv[2][9];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4* gs = reinterpret_cast<const float4*>(Gs + (8 * j + 2 * t + u) * 12);
            const float4 v0 = gs[0], v1 = gs[1], v2 = gs[2];
            gv[u][0] = v0.x; gv[u][1] = v0.y; gv[u][2] = v0.z; gv[u][3] = v0.w;
            gv[u][4] = v1.x; gv[u][5] = v1.y; gv[u][6] = v1.z; gv[u][7] = v1.w;
            gv[u][8] = v2.x;
          }
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int jj = 2 * m + half;
                float pre = br[jj];
#pragma unroll
                for (int d = 0; d < 9; ++d) pre = fmaf(gv[u][d], wr[jj][d], pre);
                dH[m][j][half * 2 + u] *= act_grad_rt<ACT>(pre, a.act);
              }
        }
        // accA[k, d] += dpre[k, n] G[n, d]: dpre (C layout) -> A fragments, G from Gb[n][16] (.trans)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t gb[4];
          ldmatrix_x4_trans(gb[0], gb[1], gb[2], gb[3], smem_u32(Gb + (ks * 16 + (mid & 1) * 8 + mr) * GB + (mid >> 1) * 8));
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            uint32_t afr[4];
            afr[0] = pack_bf16(dH[m][2 * ks][0], dH[m][2 * ks][1]);
            afr[1] = pack_bf16(dH[m][2 * ks][2], dH[m][2 * ks][3]);
            afr[2] = pack_bf16(dH[m][2 * ks + 1][0], dH[m][2 * ks + 1][1]);
            afr[3] = pack_bf16(dH[m][2 * ks + 1][2], dH[m][2 * ks + 1][3]);
            mma_bf16(accA[m][0], afr, gb[0], gb[1]);
            mma_bf16(accA[m][1], afr, gb[2], gb[3]);
          }
        }
        __syncwarp();
      }
    }
  }
  // per-warp partial [d (16)][k (32)]
  float* p = a.partials + gw * 512;
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

// one warp per output element: lanes stride over the per-warp partials, fixed-order shuffle reduce
__global__ void k_edge_tc_reduce(const float* __restrict__ partials, int n_partials, float* __restrict__ d_axes,
                                 float* __restrict__ d_bias) {
  const int lane = threadIdx.x & 31;
  const int out = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (out >= 320) return;
  float s = 0.0f;
  for (int p = lane; p < n_partials; p += 32) s += partials[(int64_t)p * 512 + out];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (out < 288) {
      if (d_axes) d_axes[out] = s;
    } else if (d_bias) {
      d_bias[out - 288] = s;
    }
  }
}

static int edge_tc_warps(int64_t n_out) {
  int64_t blocks = (n_out + AGG_WARPS - 1) / AGG_WARPS;
  const int64_t cap = (int64_t)num_sms() * 3;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks * AGG_WARPS;
}

template <int CB, int FR>
static int launch_edge_cfg(const TcEdgeArgs& a, int n_warps, cudaStream_t st) {
  constexpr int WARP_BYTES = 32 * (CB + 8) * 2 + CB * 40 * 2 + 32 * 12 * 4 + 32 * 24 * 2;
  const size_t smem = 32 * 12 * 4 + AGG_WARPS * WARP_BYTES;
  if (a.act == 2) {
    auto kern = k_edge_tc<CB, FR, 2>;
    SE3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_warps / AGG_WARPS, AGG_WARPS * 32, smem, st>>>(a);
  } else {
    auto kern = k_edge_tc<CB, FR, -1>;
    SE3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_warps / AGG_WARPS, AGG_WARPS * 32, smem, st>>>(a);
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

static int launch_edge_tc(const TcEdgeArgs& a, int f_out, int n_warps, float* dA, float* dB, cudaStream_t st) {
  int rc = SE3_OK;
  const bool small = a.c <= 16;
  switch (f_out) {
    case 1: rc = small ? launch_edge_cfg<16, 1>(a, n_warps, st) : launch_edge_cfg<32, 1>(a, n_warps, st); break;
    case 2: rc = small ? launch_edge_cfg<16, 2>(a, n_warps, st) : launch_edge_cfg<32, 2>(a, n_warps, st); break;
    case 3: rc = small ? launch_edge_cfg<16, 3>(a, n_warps, st) : launch_edge_cfg<32, 3>(a, n_warps, st); break;
    case 4: rc = small ? launch_edge_cfg<16, 4>(a, n_warps, st) : launch_edge_cfg<32, 4>(a, n_warps, st); break;
    default: set_error("launch_edge_tc: unsupported frame count"); return SE3_EINVAL;
  }
  if (rc) return rc;
  k_edge_tc_reduce<<<40, 256, 0, st>>>(a.partials, n_warps, dA, dB);
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
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Bs = As + ST * A_ROWS * A_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  const int k_begin = blockIdx.z * kchunk;
  const int k_end = min(K, k_begin + kchunk);
  const int ktiles = (k_end - k_begin + BK - 1) / BK;
  const int mid = lane >> 3, mr = lane & 7, g = lane >> 2, t = lane & 3;

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
  SE3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (splits > 1) {
    kern<<<grid, 256, smem, st>>>((int)m, (int)n, (int)k, alpha, a, lda, b, ldb, partials, n, kchunk, m * n);
    SE3_LAUNCH_CHECK();
    splitk_reduce_launch(partials, splits, m * n, alpha, reinterpret_cast<float*>(c), st);
    SE3_LAUNCH_CHECK();
  } else {
    kern<<<grid, 256, smem, st>>>((int)m, (int)n, (int)k, alpha, a, lda, b, ldb, c, ldc, kchunk, 0);
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

int gemm_tn(int64_t m, int64_t n, int64_t k, float alpha, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
            int64_t ldb, void* c, int64_t ldc, bool out_bf16, int impl, cudaStream_t st) {
  if (impl == 0) impl = gemm_impl_env();
  const bool can = tcgen05_gemm_supported(m, n, k, lda, ldb);
  if (impl == 2 && !can) {
    set_error("gemm_tn: shape not supported by the tcgen05 kernel");
    return SE3_EINVAL;
  }
  if (impl != 1 && can) return launch_gemm_tcgen05(m, n, k, alpha, a, lda, b, ldb, c, ldc, out_bf16, st);
  if (out_bf16) return launch_gemm_cfg<true, true, true>(m, n, k, alpha, a, lda, b, ldb, c, ldc, 1, nullptr, st);
  return launch_gemm_cfg<true, true, false>(m, n, k, alpha, a, lda, b, ldb, c, ldc, 1, nullptr, st);
}

// --------------------------------------------------------------------------------------------
// conversions
// --------------------------------------------------------------------------------------------
// wt[o][ck] = w[ck][o]  (bf16): K-major B operand of the forward projection
__global__ void k_transpose_w_bf16(const float* __restrict__ w, int64_t ck, int c_out, __nv_bfloat16* __restrict__ wt) {
  const int64_t total = ck * c_out;
  for (int64_t tix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; tix < total; tix += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = tix / ck, r = tix - o * ck;
    wt[tix] = __float2bfloat16(w[r * c_out + o]);
  }
}

__global__ void k_f32_to_bf16(const float* __restrict__ in, int64_t n, __nv_bfloat16* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}
// wp[c][o][k] = w[c][k][o]  (bf16)
__global__ void k_permute_w_bf16(const float* __restrict__ w, int c_in, int k, int c_out, __nv_bfloat16* __restrict__ wp) {
  const int64_t total = (int64_t)c_in * k * c_out;
  for (int64_t tix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; tix < total; tix += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(tix % k);
    const int64_t co = tix / k;
    const int o = (int)(co % c_out);
    const int64_t c = co / c_out;
    wp[tix] = __float2bfloat16(w[(c * k + kk) * c_out + o]);
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
  if (d->c_out % 8 != 0) {
    set_error("precision 1 needs c_out to be a multiple of 8 (got %d)", d->c_out);
    return SE3_EINVAL;
  }
  return SE3_OK;
}

size_t conv_tc_saved_bytes(const se3_conv_desc* d) {
  return align_up((size_t)d->n_out * d->f_out * d->c_in * d->k * 2) + 256;  // T in bf16
}
size_t conv_tc_fwd_workspace_bytes(const se3_conv_desc* d) {
  return align_up((size_t)d->c_in * d->k * d->c_out * 2) + 256;  // W in bf16
}
size_t conv_tc_bwd_workspace_bytes(const se3_conv_desc* d) {
  const int64_t R = d->n_out * d->f_out, ck = (int64_t)d->c_in * d->k, Nf = d->n_in * d->f_in;
  size_t b = 0;
  b += align_up((size_t)ck * d->c_out * 2) * 2;                                   // Wb, Wpb
  b += align_up((size_t)R * d->c_out * 2);                                        // dy bf16
  b += align_up((size_t)R * ck * 2);                                              // dT bf16
  b += align_up((size_t)Nf * d->c_out * d->k * 2);                                // U bf16
  b += align_up((size_t)splits_for_tc(ck, d->c_out, R) * ck * d->c_out * 4);      // dW partials
  b += align_up((size_t)edge_tc_warps(d->n_out) * 512 * 4);                       // basis-gradient partials
  return b + 256;
}

int conv_tc_fwd(const se3_conv_desc* d, const float* x, float* y, void* saved, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  if (int rc = check_tc(d)) return rc;
  SE3_CHECK_ARG(saved && ws, "precision 1 needs the saved buffer and a workspace");
  const int64_t R = d->n_out * d->f_out, ck = (int64_t)d->c_in * d->k;
  Arena ar(ws, ws_bytes);
  __nv_bfloat16* Wb = ar.take<__nv_bfloat16>(ck * d->c_out);
  if (!ar.ok()) { set_error("conv_tc_fwd: workspace too small"); return SE3_EWORKSPACE; }
  __nv_bfloat16* T = reinterpret_cast<__nv_bfloat16*>(saved);
  k_transpose_w_bf16<<<blocks_for(ck * d->c_out), 256, 0, st>>>(d->conv_weights, ck, d->c_out, Wb);
  SE3_LAUNCH_CHECK();
  TcAggArgs a;
  a.row_ends = d->row_ends; a.nbr = d->col_src;
  a.pts_row = d->pts_out; a.frm_row = d->frames_out;
  a.pts_g = d->pts_in; a.frm_g = d->frames_in; a.f_g = d->f_in;
  a.feat = x; a.c = d->c_in; a.w9 = d->proj_axes; a.bias = d->proj_biases; a.norm = d->norm_neigh_dist;
  a.act = d->act; a.out = T; a.n_rows = d->n_out;
  if (int rc = launch_agg_tc<false>(a, d->f_out, st)) return rc;
  // y[r,o] = s * sum_(c,k) T[r,(c,k)] Wt[o,(c,k)]   (tcgen05 / TMEM)
  return gemm_tn(R, d->c_out, ck, d->out_scale, T, ck, Wb, ck, y, d->c_out, false, 0, st);
}

int conv_tc_bwd(const se3_conv_desc* d, const float* x, const float* dy, const void* saved, float* dx, float* dW,
                float* dA, float* dB, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (int rc = check_tc(d)) return rc;
  SE3_CHECK_ARG(saved && ws, "precision 1 backward needs the forward's saved buffer and a workspace");
  SE3_CHECK_ARG(!dx || (d->t_row_ends && d->t_edge && d->t_dst), "dx needs the transposed CSR");
  const int64_t R = d->n_out * d->f_out, ck = (int64_t)d->c_in * d->k, Nf = d->n_in * d->f_in;
  const int64_t ok = (int64_t)d->c_out * d->k;
  Arena ar(ws, ws_bytes);
  __nv_bfloat16* Wb = ar.take<__nv_bfloat16>(ck * d->c_out);
  __nv_bfloat16* Wpb = ar.take<__nv_bfloat16>(ck * d->c_out);
  __nv_bfloat16* dyb = ar.take<__nv_bfloat16>(R * d->c_out);
  __nv_bfloat16* dTb = ar.take<__nv_bfloat16>(R * ck);
  __nv_bfloat16* U = ar.take<__nv_bfloat16>(Nf * ok);
  const int splits = splits_for_tc(ck, d->c_out, R);
  float* dWp = ar.take<float>((size_t)splits * ck * d->c_out);
  const int n_warps = edge_tc_warps(d->n_out);
  float* eg = ar.take<float>((size_t)n_warps * 512);
  if (!ar.ok()) { set_error("conv_tc_bwd: workspace too small"); return SE3_EWORKSPACE; }
  const __nv_bfloat16* T = reinterpret_cast<const __nv_bfloat16*>(saved);
  k_f32_to_bf16<<<blocks_for(R * d->c_out), 256, 0, st>>>(dy, R * d->c_out, dyb);
  SE3_LAUNCH_CHECK();
  if (dW) {
    // dW[(c,k), o] = s * sum_r T[r,(c,k)] dy[r,o]
    if (int rc = launch_gemm_cfg<false, false, false>(ck, d->c_out, R, d->out_scale, T, ck, dyb, d->c_out, dW, d->c_out,
                                                      splits, dWp, st))
      return rc;
  }
  if (dA || dB) {
    k_f32_to_bf16<<<blocks_for(ck * d->c_out), 256, 0, st>>>(d->conv_weights, ck * d->c_out, Wb);
    SE3_LAUNCH_CHECK();
    // dT[r,(c,k)] = s * sum_o dy[r,o] W[(c,k),o]
    if (int rc = gemm_tn(R, ck, d->c_out, d->out_scale, dyb, d->c_out, Wb, d->c_out, dTb, ck, true, 0, st)) return rc;
    TcEdgeArgs g;
    g.row_ends = d->row_ends; g.col_src = d->col_src;
    g.pts_out = d->pts_out; g.frm_out = d->frames_out;
    g.pts_in = d->pts_in; g.frm_in = d->frames_in; g.f_in = d->f_in;
    g.x = x; g.c = d->c_in; g.w9 = d->proj_axes; g.bias = d->proj_biases; g.norm = d->norm_neigh_dist;
    g.act = d->act; g.dT = dTb; g.n_out = d->n_out; g.partials = eg;
    if (int rc = launch_edge_tc(g, d->f_out, n_warps, dA, dB, st)) return rc;
  }
  if (dx) {
    TcAggArgs a;
    a.row_ends = d->t_row_ends; a.nbr = d->t_dst;
    a.pts_row = d->pts_in; a.frm_row = d->frames_in;
    a.pts_g = d->pts_out; a.frm_g = d->frames_out; a.f_g = d->f_out;
    a.feat = dy; a.c = d->c_out; a.w9 = d->proj_axes; a.bias = d->proj_biases; a.norm = d->norm_neigh_dist;
    a.act = d->act; a.out = U; a.n_rows = d->n_in;
    if (int rc = launch_agg_tc<true>(a, d->f_in, st)) return rc;
    k_permute_w_bf16<<<blocks_for(ck * d->c_out), 256, 0, st>>>(d->conv_weights, d->c_in, d->k, d->c_out, Wpb);
    SE3_LAUNCH_CHECK();
    // dx[n,c] = s * sum_(o,k) U[n,(o,k)] Wp[c,(o,k)]
    if (int rc = gemm_tn(Nf, d->c_in, ok, d->out_scale, U, ok, Wpb, ok, dx, d->c_in, false, 0, st)) return rc;
  }
  return SE3_OK;
}

}  // namespace se3

// C = alpha * A[M,K] . B[N,K]^T, bf16 operands (K-major), fp32 or bf16 output.  impl: 0 auto, 1 mma.sync, 2 tcgen05.
extern "C" int se3_gemm_bf16_tn(const void* a, const void* b, int64_t m, int64_t n, int64_t k, float alpha, void* c,
                                int32_t c_is_bf16, int32_t impl, se3_stream_t stream) {
  using namespace se3;
  SE3_CHECK_ARG(m >= 0 && n >= 1 && k >= 8 && (k % 8) == 0, "bad sizes (k must be a multiple of 8)");
  SE3_CHECK_ARG(impl >= 0 && impl <= 2, "impl must be 0, 1 or 2");
  if (m == 0) return SE3_OK;
  SE3_CHECK_ARG(a && b && c, "null pointer");
  return gemm_tn(m, n, k, alpha, reinterpret_cast<const __nv_bfloat16*>(a), k, reinterpret_cast<const __nv_bfloat16*>(b), k,
                 c, n, c_is_bf16 != 0, impl, as_stream(stream));
}
