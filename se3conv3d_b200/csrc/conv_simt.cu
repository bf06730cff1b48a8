// fp32 CUDA-core path of the fused PNEConvLayerRotEquiv (precision 0, "exactness mode") and the
// legacy feat_basis_proj ops.  See conv_simt.cuh / include/se3conv3d_b200.h for the reference
// citations.  Everything here is deterministic (no floating-point atomics) except the legacy
// feat_basis_proj_grad feature gradient, which keeps the reference's atomic formulation because its
// signature carries no transposed CSR.
#include "conv_simt.cuh"

namespace se3 {

// --------------------------------------------------------------------------------------------
// geometry: lanes 0..8 of a warp each produce one component of g[e,a,b]
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float geom_component(int lane, const float* __restrict__ Ro, const float* __restrict__ Ri,
                                                float dx, float dy, float dz) {
  float g = 0.0f;
  if (lane < 3) {
    // u = d^T R_out  (row vector times matrix, RotationFunctions.py:659-661)
    g = dx * __ldg(Ro + lane) + dy * __ldg(Ro + 3 + lane) + dz * __ldg(Ro + 6 + lane);
  } else if (lane < 9) {
    // rows 0,1 of R_out^T R_in  (RotationFunctions.py:589-592, 252)
    const int m = (lane - 3) / 3, n = (lane - 3) % 3;
    g = __ldg(Ro + m) * __ldg(Ri + n) + __ldg(Ro + 3 + m) * __ldg(Ri + 3 + n) + __ldg(Ro + 6 + m) * __ldg(Ri + 6 + n);
  }
  return g;
}

template <bool TR>
__global__ void __launch_bounds__(128) k_aggregate_f32(const AggArgs a, const int ncb) {
  __shared__ __align__(16) float xs[4][2][32];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // one warp per (row point, row frame, 32-channel block)
  const int per_row = ncb * a.f_row;
  const int64_t total = a.n_rows * per_row;
  float wk[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) wk[d] = a.w9[d * 32 + lane];
  const float bk = a.bias[lane];
  int it = 0;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += nwarps) {
    const int64_t rp = w / per_row;
    const int sub = (int)(w - rp * per_row);
    const int fr = sub / ncb, cb = sub - fr * ncb;
    const int lo = rp > 0 ? a.row_ends[rp - 1] : 0;
    const int hi = a.row_ends[rp];
    const float prx = a.pts_row[3 * rp], pry = a.pts_row[3 * rp + 1], prz = a.pts_row[3 * rp + 2];
    const int ch = cb * 32 + lane;
    {
      const float* Frow = a.frm_row + (rp * a.f_row + fr) * 9;
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
      for (int e = lo; e < hi; ++e) {
        const int64_t q = a.nbr[e];
        const float pgx = __ldg(a.pts_g + 3 * q), pgy = __ldg(a.pts_g + 3 * q + 1), pgz = __ldg(a.pts_g + 3 * q + 2);
        const float dx = (TR ? (prx - pgx) : (pgx - prx)) * a.norm;
        const float dy = (TR ? (pry - pgy) : (pgy - pry)) * a.norm;
        const float dz = (TR ? (prz - pgz) : (pgz - prz)) * a.norm;
        for (int fg = 0; fg < a.f_g; ++fg) {
          const float* Fg = a.frm_g + (q * a.f_g + fg) * 9;
          const float g = geom_component(lane, TR ? Fg : Frow, TR ? Frow : Fg, dx, dy, dz);
          float pre = bk;
#pragma unroll
          for (int d = 0; d < 9; ++d) pre = fmaf(__shfl_sync(0xffffffffu, g, d), wk[d], pre);
          const float h = pne_act(pre, a.act);
          const float xv = ch < a.c ? __ldg(a.feat + (q * a.f_g + fg) * a.c + ch) : 0.0f;
          float* buf = xs[wib][it & 1];
          buf[lane] = xv;
          __syncwarp();
          const float4* b4 = reinterpret_cast<const float4*>(buf);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = b4[i];
            acc[4 * i + 0] = fmaf(v.x, h, acc[4 * i + 0]);
            acc[4 * i + 1] = fmaf(v.y, h, acc[4 * i + 1]);
            acc[4 * i + 2] = fmaf(v.z, h, acc[4 * i + 2]);
            acc[4 * i + 3] = fmaf(v.w, h, acc[4 * i + 3]);
          }
          ++it;
        }
      }
      float* o = a.out + ((rp * a.f_row + fr) * a.c + (int64_t)cb * 32) * 32 + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (cb * 32 + i < a.c) o[i * 32] = acc[i];
    }
  }
}

int launch_aggregate_f32(const AggArgs& a, bool transposed, cudaStream_t st) {
  if (a.n_rows == 0) return SE3_OK;
  const int ncb = (a.c + 31) / 32;
  const int64_t warps = a.n_rows * ncb * a.f_row;
  int64_t blocks = (warps + 3) / 4;
  const int64_t cap = (int64_t)num_sms() * 64;
  if (blocks > cap) blocks = cap;
  if (transposed)
    k_aggregate_f32<true><<<(int)blocks, 128, 0, st>>>(a, ncb);
  else
    k_aggregate_f32<false><<<(int)blocks, 128, 0, st>>>(a, ncb);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// --------------------------------------------------------------------------------------------
// SGEMM (64x64x16 tiles, 4x4 micro-tiles), optional deterministic split-K
// --------------------------------------------------------------------------------------------
template <bool AT, bool BT>
__global__ void __launch_bounds__(256) k_sgemm(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t lda,
                                               const float* __restrict__ B, int64_t ldb, float* __restrict__ C,
                                               int64_t ldc, int kchunk, int64_t partial_stride) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int t = threadIdx.x;
  const int bm = blockIdx.y * 64, bn = blockIdx.x * 64;
  const int k0 = blockIdx.z * kchunk;
  const int k1 = min(K, k0 + kchunk);
  const int tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int kk = k0; kk < k1; kk += 16) {
    if (!AT) {
      const int m = t >> 2, kq = (t & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = kk + kq + i;
        As[kq + i][m] = (bm + m < M && k < k1) ? A[(int64_t)(bm + m) * lda + k] : 0.0f;
      }
    } else {
      const int k = t >> 4, mq = (t & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = bm + mq + i;
        As[k][mq + i] = (kk + k < k1 && m < M) ? A[(int64_t)(kk + k) * lda + m] : 0.0f;
      }
    }
    if (!BT) {
      const int k = t >> 4, nq = (t & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = bn + nq + i;
        Bs[k][nq + i] = (kk + k < k1 && n < N) ? B[(int64_t)(kk + k) * ldb + n] : 0.0f;
      }
    } else {
      const int n = t >> 2, kq = (t & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = kk + kq + i;
        Bs[kq + i][n] = (bn + n < N && k < k1) ? B[(int64_t)(bn + n) * ldb + k] : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Cz = C + (int64_t)blockIdx.z * partial_stride;
  const float sc = partial_stride ? 1.0f : alpha;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = bm + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = bn + tx * 4 + j;
      if (n < N) Cz[(int64_t)m * ldc + n] = sc * acc[i][j];
    }
  }
}

__global__ void k_splitk_reduce(const float* __restrict__ partials, int splits, int64_t mn, float alpha,
                                float* __restrict__ c) {
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mn; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int z = 0; z < splits; ++z) s += partials[(int64_t)z * mn + i];
    c[i] = alpha * s;
  }
}
// eight split loads in flight per thread; the sums stay in split order (same bits as the plain loop)
__global__ void __launch_bounds__(256) k_splitk_reduce8(const float* __restrict__ partials, int splits, int64_t mn, float alpha,
                                                        float* __restrict__ c) {
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mn; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    int z = 0;
    for (; z + 8 <= splits; z += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(partials + (int64_t)(z + u) * mn + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; z < splits; ++z) s += __ldcs(partials + (int64_t)z * mn + i);
    c[i] = alpha * s;
  }
}

void splitk_reduce_launch(const float* partials, int splits, int64_t mn, float alpha, float* c, cudaStream_t st) {
  int64_t blocks = (mn + 255) / 256;
  if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
  if (splits >= 8)
    launch_pdl(k_splitk_reduce8, dim3((unsigned)blocks), dim3(256), 0, st, partials, splits, mn, alpha, c);
  else
    launch_pdl(k_splitk_reduce, dim3((unsigned)blocks), dim3(256), 0, st, partials, splits, mn, alpha, c);
}

int launch_sgemm(bool at, bool bt, int64_t m, int64_t n, int64_t k, float alpha, const float* a, int64_t lda,
                 const float* b, int64_t ldb, float* c, int64_t ldc, int splits, float* partials, cudaStream_t st) {
  if (m == 0 || n == 0) return SE3_OK;
  if (splits < 1 || partials == nullptr) splits = 1;
  int kchunk = (int)((k + splits - 1) / splits);
  kchunk = (kchunk + 15) / 16 * 16;
  if (kchunk < 16) kchunk = 16;
  splits = (int)((k + kchunk - 1) / kchunk);
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)((n + 63) / 64), (unsigned)((m + 63) / 64), (unsigned)splits);
  float* dst = splits > 1 ? partials : c;
  const int64_t pstride = splits > 1 ? m * n : 0;
  const int64_t ld_out = splits > 1 ? n : ldc;
  if (!at && !bt)
    k_sgemm<false, false><<<grid, 256, 0, st>>>((int)m, (int)n, (int)k, alpha, a, lda, b, ldb, dst, ld_out, kchunk, pstride);
  else if (!at && bt)
    k_sgemm<false, true><<<grid, 256, 0, st>>>((int)m, (int)n, (int)k, alpha, a, lda, b, ldb, dst, ld_out, kchunk, pstride);
  else if (at && !bt)
    k_sgemm<true, false><<<grid, 256, 0, st>>>((int)m, (int)n, (int)k, alpha, a, lda, b, ldb, dst, ld_out, kchunk, pstride);
  else
    k_sgemm<true, true><<<grid, 256, 0, st>>>((int)m, (int)n, (int)k, alpha, a, lda, b, ldb, dst, ld_out, kchunk, pstride);
  SE3_LAUNCH_CHECK();
  if (splits > 1) {
    if (ldc != n) {
      set_error("launch_sgemm: split-K needs a dense C");
      return SE3_EINVAL;
    }
    int64_t blocks = (m * n + 255) / 256;
    if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
    k_splitk_reduce<<<(int)blocks, 256, 0, st>>>(partials, splits, m * n, alpha, c);
    SE3_LAUNCH_CHECK();
  }
  return SE3_OK;
}

// --------------------------------------------------------------------------------------------
// gradient of the basis parameters (proj_axes_, proj_biases_), by output row, deterministic
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_edge_grad_f32(const EdgeGradArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* xs = smem + (size_t)wib * a.c;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float wk[9], acc9[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    wk[d] = a.w9[d * 32 + lane];
    acc9[d] = 0.0f;
  }
  const float bk = a.bias[lane];
  float accb = 0.0f;
  for (int64_t rp = gw; rp < a.n_out; rp += nwarps) {
    const int lo = rp > 0 ? a.row_ends[rp - 1] : 0;
    const int hi = a.row_ends[rp];
    const float prx = a.pts_out[3 * rp], pry = a.pts_out[3 * rp + 1], prz = a.pts_out[3 * rp + 2];
    for (int fr = 0; fr < a.f_out; ++fr) {
      const float* Fo = a.frm_out + (rp * a.f_out + fr) * 9;
      const float* dTrow = a.dT + (rp * a.f_out + fr) * (int64_t)a.c * 32;
      // up to 32 channels: this lane's column of the dT tile lives in registers for the whole row
      float dTr[32];
      const bool small_c = a.c <= 32;
      if (small_c) {
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) dTr[cc] = cc < a.c ? __ldg(dTrow + cc * 32 + lane) : 0.0f;
      }
      for (int e = lo; e < hi; ++e) {
        const int64_t q = a.col_src[e];
        const float dx = (__ldg(a.pts_in + 3 * q) - prx) * a.norm;
        const float dy = (__ldg(a.pts_in + 3 * q + 1) - pry) * a.norm;
        const float dz = (__ldg(a.pts_in + 3 * q + 2) - prz) * a.norm;
        for (int fg = 0; fg < a.f_in; ++fg) {
          const float g = geom_component(lane, Fo, a.frm_in + (q * a.f_in + fg) * 9, dx, dy, dz);
          float gd[9];
          float pre = bk;
#pragma unroll
          for (int d = 0; d < 9; ++d) {
            gd[d] = __shfl_sync(0xffffffffu, g, d);
            pre = fmaf(gd[d], wk[d], pre);
          }
          const float* xrow = a.x + (q * a.f_in + fg) * a.c;
          float dH = 0.0f;
          if (small_c) {
            const float xv = lane < a.c ? __ldg(xrow + lane) : 0.0f;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) dH = fmaf(dTr[cc], __shfl_sync(0xffffffffu, xv, cc), dH);
          } else {
            for (int cc = lane; cc < a.c; cc += 32) xs[cc] = __ldg(xrow + cc);
            __syncwarp();
            for (int cc = 0; cc < a.c; ++cc) dH = fmaf(__ldg(dTrow + cc * 32 + lane), xs[cc], dH);
            __syncwarp();
          }
          const float dpre = dH * pne_act_grad(pre, a.act);
#pragma unroll
          for (int d = 0; d < 9; ++d) acc9[d] = fmaf(gd[d], dpre, acc9[d]);
          accb += dpre;
        }
      }
    }
  }
  float* p = a.partials + gw * 320;
#pragma unroll
  for (int d = 0; d < 9; ++d) p[d * 32 + lane] = acc9[d];
  p[9 * 32 + lane] = accb;
}

// Ordered reduction of the per-warp partials [n_partials][320]: block b owns outputs 32b..32b+31; warp w sums
// partials w, w+32, ... (lane = output, coalesced 128-byte rows), then the 32 warp sums are added in warp order.
__global__ void __launch_bounds__(1024) k_edge_grad_reduce(const float* __restrict__ partials, int n_partials,
                                                           float* __restrict__ d_axes, float* __restrict__ d_bias) {
  __shared__ float sm[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + lane;
  float s = 0.0f;
  for (int p = w; p < n_partials; p += 32) s += partials[(int64_t)p * 320 + t];
  sm[w][lane] = s;
  __syncthreads();
  if (w == 0) {
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) tot += sm[i][lane];
    if (t < 288) {
      if (d_axes) d_axes[t] = tot;
    } else if (d_bias) {
      d_bias[t - 288] = tot;
    }
  }
}

int edge_grad_partials(int64_t n_out) {
  int64_t blocks = (n_out + 3) / 4;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks * 4;
}

int launch_edge_grad_f32(const EdgeGradArgs& a, float* d_axes, float* d_bias, cudaStream_t st) {
  const int blocks = a.n_partials / 4;
  const size_t smem = (size_t)4 * a.c * sizeof(float);
  k_edge_grad_f32<<<blocks, 128, smem, st>>>(a);
  SE3_LAUNCH_CHECK();
  k_edge_grad_reduce<<<10, 1024, 0, st>>>(a.partials, a.n_partials, d_axes, d_bias);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// --------------------------------------------------------------------------------------------
// legacy ops: feat_basis_proj / feat_basis_proj_grad
// --------------------------------------------------------------------------------------------
__global__ void k_feat_basis_proj(const float* __restrict__ basis, const float* __restrict__ feats,
                                  const int2* __restrict__ nb, const int* __restrict__ ends, int64_t m, int c, int k,
                                  float* __restrict__ out) {
  // one thread per output element T[row, ch, kk]; threads of a warp share (row, ch) when k >= 32
  const int64_t total = m * c * k;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(t % k);
    const int64_t rc = t / k;
    const int ch = (int)(rc % c);
    const int64_t row = rc / c;
    const int lo = row > 0 ? ends[row - 1] : 0, hi = ends[row];
    float acc = 0.0f;
    for (int e = lo; e < hi; ++e) acc = fmaf(__ldg(feats + (int64_t)nb[e].y * c + ch), __ldg(basis + (int64_t)e * k + kk), acc);
    out[t] = acc;
  }
}

__global__ void k_feat_basis_grad_basis(const float* __restrict__ feats, const int2* __restrict__ nb,
                                        const float* __restrict__ grads, int64_t n_edges, int c, int k,
                                        float* __restrict__ basis_grads) {
  const int64_t total = n_edges * k;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(t % k);
    const int64_t e = t / k;
    const int2 p = nb[e];
    const float* g = grads + (int64_t)p.x * c * k + kk;
    const float* f = feats + (int64_t)p.y * c;
    float acc = 0.0f;
    for (int ch = 0; ch < c; ++ch) acc = fmaf(__ldg(g + (int64_t)ch * k), __ldg(f + ch), acc);
    basis_grads[t] = acc;
  }
}

__global__ void k_feat_basis_grad_feat(const float* __restrict__ basis, const int2* __restrict__ nb,
                                       const float* __restrict__ grads, int64_t n_edges, int c, int k,
                                       float* __restrict__ feat_grads) {
  const int64_t total = n_edges * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(t % c);
    const int64_t e = t / c;
    const int2 p = nb[e];
    const float* g = grads + ((int64_t)p.x * c + ch) * k;
    const float* b = basis + e * k;
    float acc = 0.0f;
    for (int kk = 0; kk < k; ++kk) acc = fmaf(__ldg(g + kk), __ldg(b + kk), acc);
    atomicAdd(feat_grads + (int64_t)p.y * c + ch, acc);
  }
}

__global__ void k_permute_w(const float* __restrict__ w, int c_in, int k, int c_out, float* __restrict__ wp) {
  // wp[c][o][kk] = w[c][kk][o]
  const int64_t total = (int64_t)c_in * k * c_out;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(t % k);
    const int64_t co = t / k;
    const int o = (int)(co % c_out);
    const int64_t c = co / c_out;
    wp[t] = w[(c * k + kk) * c_out + o];
  }
}

static inline int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int launch_permute_w(const float* w, int c_in, int k, int c_out, float* wp, cudaStream_t st) {
  k_permute_w<<<blocks_for((int64_t)c_in * k * c_out), 256, 0, st>>>(w, c_in, k, c_out, wp);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

}  // namespace se3

using namespace se3;

extern "C" int se3_feat_basis_proj(const float* basis, const float* feats, const int32_t* neighbors,
                                   const int32_t* ends, int64_t n_edges, int64_t m, int32_t c, int32_t k, float* out,
                                   se3_stream_t stream) {
  (void)n_edges;
  SE3_CHECK_ARG(k == 8 || k == 16 || k == 32 || k == 64, "K must be 8, 16, 32 or 64");
  SE3_CHECK_ARG(c >= 1 && m >= 0, "bad sizes");
  if (m == 0) return SE3_OK;
  SE3_CHECK_ARG(basis && feats && neighbors && ends && out, "null pointer");
  k_feat_basis_proj<<<blocks_for(m * c * k), 256, 0, as_stream(stream)>>>(
      basis, feats, reinterpret_cast<const int2*>(neighbors), ends, m, c, k, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_feat_basis_proj_grad(const float* basis, const float* feats, const int32_t* neighbors,
                                        const int32_t* ends, const float* grads, int64_t n_edges, int64_t m,
                                        int64_t n_feat_rows, int32_t c, int32_t k, float* feat_grads,
                                        float* basis_grads, se3_stream_t stream) {
  (void)ends;
  (void)m;
  SE3_CHECK_ARG(k == 8 || k == 16 || k == 32 || k == 64, "K must be 8, 16, 32 or 64");
  SE3_CHECK_ARG(c >= 1 && n_edges >= 0 && n_feat_rows >= 0, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (feat_grads && n_feat_rows > 0) SE3_CUDA(cudaMemsetAsync(feat_grads, 0, n_feat_rows * c * sizeof(float), st));
  if (n_edges == 0) return SE3_OK;
  SE3_CHECK_ARG(basis && feats && neighbors && grads, "null pointer");
  const int2* nb = reinterpret_cast<const int2*>(neighbors);
  if (basis_grads) {
    k_feat_basis_grad_basis<<<blocks_for(n_edges * k), 256, 0, st>>>(feats, nb, grads, n_edges, c, k, basis_grads);
    SE3_LAUNCH_CHECK();
  }
  if (feat_grads) {
    k_feat_basis_grad_feat<<<blocks_for(n_edges * c), 256, 0, st>>>(basis, nb, grads, n_edges, c, k, feat_grads);
    SE3_LAUNCH_CHECK();
  }
  return SE3_OK;
}
