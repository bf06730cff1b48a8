// The elementwise / pooling glue around the convolution as single kernels:
//   se3_gamma_skip_fwd / _bwd   out = drop_path(x * gamma) + y                 layers/SkipConnection.py:31-43 with
//                               layers/DropPathPC.py:23-50 (per batch item keep mask, 1 / keep_prob rescale)
//   se3_frame_pool_fwd / _bwd   pooling of the F per-frame rows of a point     pc/PointcloudRotEquiv.py:224-251
//   se3_batch_pool_fwd / _bwd   pooling of the rows of a batch item (sorted ids) pc/PointcloudRotEquiv.py:195-222, 253-275;
//                               its backward with mode "gather" is global_upsample (:277-286)
// All reductions are ordered (no atomics): bit-identical from run to run.
#include "common.cuh"

namespace se3 {
namespace {

constexpr int kRowsPerBlock = 64;   // rows per block of the gamma-skip backward: ~1 wave of blocks at dfaust sizes

// out[r,c] = x[r,c] * gamma[c] * s(r) + y[r,c],  s(r) = row_scale[row_batch[r] / rows_per_id] (1 when row_scale is null)
__global__ void __launch_bounds__(256) k_gamma_skip_fwd(const float* __restrict__ x, const float* __restrict__ y,
                                                        const float* __restrict__ gamma, const float* __restrict__ row_scale,
                                                        const int32_t* __restrict__ row_batch, int frames, int64_t rows, int c,
                                                        float* __restrict__ out) {
  const int64_t total = rows * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int ch = (int)(i - r * c);
    const float s = row_scale ? row_scale[row_batch[r / frames]] : 1.0f;
    out[i] = fmaf(x[i] * gamma[ch], s, y[i]);
  }
}

// dx[r,c] = dy[r,c] * gamma[c] * s(r);  partial[b, c] = sum over the rows of block b of dy * x * s
__global__ void __launch_bounds__(256) k_gamma_skip_bwd(const float* __restrict__ dy, const float* __restrict__ x,
                                                        const float* __restrict__ gamma, const float* __restrict__ row_scale,
                                                        const int32_t* __restrict__ row_batch, int frames, int64_t rows, int c,
                                                        float* __restrict__ dx, float* __restrict__ partial) {
  extern __shared__ float red[];   // [256 / cpt][c] partial sums of the thread groups
  const int64_t r0 = (int64_t)blockIdx.x * kRowsPerBlock;
  const int64_t r1 = min(rows, r0 + kRowsPerBlock);
  // thread t owns channel t % c of rows t / c, t / c + 256 / c ... (c <= 256) or loops over channels (c > 256)
  const int groups = c <= 256 ? 256 / c : 1;
  const int g = c <= 256 ? threadIdx.x / c : 0;
  for (int ch = c <= 256 ? threadIdx.x % c : threadIdx.x; ch < c; ch += 256) {
    float acc = 0.0f;
    if (g < groups) {
      const float gm = gamma[ch];
#pragma unroll 4
      for (int64_t r = r0 + g; r < r1; r += groups) {
        const float s = row_scale ? row_scale[row_batch[r / frames]] : 1.0f;
        const float d = dy[r * c + ch];
        if (dx) dx[r * c + ch] = d * gm * s;
        acc = fmaf(d * s, x[r * c + ch], acc);
      }
    }
    if (c <= 256) {
      if (g < groups) red[g * c + ch] = acc;
      __syncthreads();
      if (g == 0) {
        float t = 0.0f;
        for (int q = 0; q < groups; ++q) t += red[q * c + ch];   // fixed order
        partial[(int64_t)blockIdx.x * c + ch] = t;
      }
    } else {
      partial[(int64_t)blockIdx.x * c + ch] = acc;
    }
  }
}
// ordered two-level sum of the per-block partials: block = 32 channels, warp w adds partials w, w + 32, ... (128-byte rows),
// then the 32 warp sums are added in warp order
__global__ void __launch_bounds__(1024) k_partial_reduce(const float* __restrict__ partial, int nb, int c, float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;
  float t = 0.0f;
  if (ch < c)
    for (int b = w; b < nb; b += 32) t += partial[(int64_t)b * c + ch];
  sm[w][lane] = t;
  __syncthreads();
  if (w == 0 && ch < c) {
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) tot += sm[i][lane];
    out[ch] = tot;
  }
}

// mode: 0 avg, 1 sum, 2 max, 3 min over the f consecutive rows of a point
__global__ void __launch_bounds__(256) k_frame_pool_fwd(const float* __restrict__ x, int64_t n, int f, int c, int mode,
                                                        float* __restrict__ out) {
  const int64_t total = n * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / c;
    const int ch = (int)(i - p * c);
    const float* src = x + (p * f) * c + ch;
    float v = src[0];
    for (int a = 1; a < f; ++a) {
      const float w = src[(int64_t)a * c];
      v = mode <= 1 ? v + w : (mode == 2 ? fmaxf(v, w) : fminf(v, w));
    }
    out[i] = mode == 0 ? v / (float)f : v;
  }
}
// gradient of the frame pooling; max / min route to the first frame that attains the extremum
__global__ void __launch_bounds__(256) k_frame_pool_bwd(const float* __restrict__ dout, const float* __restrict__ x,
                                                        const float* __restrict__ out, int64_t n, int f, int c, int mode,
                                                        float* __restrict__ dx) {
  const int64_t total = n * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / c;
    const int ch = (int)(i - p * c);
    const float d = dout[i];
    bool given = false;
    for (int a = 0; a < f; ++a) {
      const int64_t j = (p * f + a) * c + ch;
      float v;
      if (mode <= 1) {
        v = mode == 0 ? d / (float)f : d;
      } else {
        const bool hit = !given && x[j] == out[i];
        given = given || hit;
        v = hit ? d : 0.0f;
      }
      dx[j] = v;
    }
  }
}

// rows of batch item b = [ends[b-1], ends[b]) (sorted ids); one block column per (item, 32 channels), ordered sums
__global__ void __launch_bounds__(256) k_batch_pool_fwd(const float* __restrict__ x, const int32_t* __restrict__ ends, int c,
                                                        int mode, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int b = blockIdx.x, ch = blockIdx.y * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  const int lo = b ? ends[b - 1] : 0, hi = ends[b];
  float v = mode <= 1 ? 0.0f : (mode == 2 ? -3.4e38f : 3.4e38f);
  if (ch < c)
    for (int r = lo + w; r < hi; r += 8) {
      const float t = x[(int64_t)r * c + ch];
      v = mode <= 1 ? v + t : (mode == 2 ? fmaxf(v, t) : fminf(v, t));
    }
  red[w][threadIdx.x & 31] = v;
  __syncthreads();
  if (w == 0 && ch < c) {
    float t = red[0][threadIdx.x];
    for (int q = 1; q < 8; ++q) t = mode <= 1 ? t + red[q][threadIdx.x] : (mode == 2 ? fmaxf(t, red[q][threadIdx.x]) : fminf(t, red[q][threadIdx.x]));
    if (mode == 0) t = hi > lo ? t / (float)(hi - lo) : 0.0f;
    if (mode >= 2 && hi == lo) t = 0.0f;
    out[(int64_t)b * c + ch] = t;
  }
}
// dx[r,c] = dout[item(r), c] (* 1 / count for avg); mode 4 = plain gather (global_upsample)
__global__ void __launch_bounds__(256) k_batch_pool_bwd(const float* __restrict__ dout, const int32_t* __restrict__ ends,
                                                        const int32_t* __restrict__ row_item, int64_t rows, int c, int mode,
                                                        float* __restrict__ dx) {
  const int64_t total = rows * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int ch = (int)(i - r * c);
    const int b = row_item[r];
    float v = dout[(int64_t)b * c + ch];
    if (mode == 0) {
      const int lo = b ? ends[b - 1] : 0;
      v /= (float)max(ends[b] - lo, 1);
    }
    dx[i] = v;
  }
}

inline int grid_for_elems(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)std::max<int64_t>(1, std::min(b, cap));
}

}  // namespace
}  // namespace se3

using namespace se3;

extern "C" size_t se3_gamma_skip_workspace_bytes(int64_t rows, int32_t c) {
  return align_up((size_t)((rows + kRowsPerBlock - 1) / kRowsPerBlock) * (size_t)c * sizeof(float)) + 256;
}

extern "C" int se3_gamma_skip_fwd(const float* x, const float* y, const float* gamma, const float* item_scale,
                                  const int32_t* point_item, int32_t frames, int64_t rows, int32_t c, float* out,
                                  se3_stream_t stream) {
  SE3_CHECK_ARG(rows >= 0 && c >= 1 && frames >= 1, "bad sizes");
  if (rows == 0) return SE3_OK;
  SE3_CHECK_ARG(x && y && gamma && out && (!item_scale || point_item), "null pointer");
  k_gamma_skip_fwd<<<grid_for_elems(rows * c), 256, 0, as_stream(stream)>>>(x, y, gamma, item_scale, point_item, frames, rows, c, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_gamma_skip_bwd(const float* dy, const float* x, const float* gamma, const float* item_scale,
                                  const int32_t* point_item, int32_t frames, int64_t rows, int32_t c, float* dx, float* dgamma,
                                  void* workspace, size_t workspace_bytes, se3_stream_t stream) {
  SE3_CHECK_ARG(rows >= 0 && c >= 1 && frames >= 1, "bad sizes");
  SE3_CHECK_ARG(dgamma != nullptr, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (rows == 0) {
    SE3_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)c * sizeof(float), st));
    return SE3_OK;
  }
  SE3_CHECK_ARG(dy && x && gamma && workspace && (!item_scale || point_item), "null pointer");
  const int nb = (int)((rows + kRowsPerBlock - 1) / kRowsPerBlock);
  if (workspace_bytes < (size_t)nb * c * sizeof(float)) {
    set_error("se3_gamma_skip_bwd: workspace too small");
    return SE3_EWORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(workspace);
  const size_t smem = c <= 256 ? (size_t)(256 / c) * c * sizeof(float) : 0;
  k_gamma_skip_bwd<<<nb, 256, smem, st>>>(dy, x, gamma, item_scale, point_item, frames, rows, c, dx, partial);
  SE3_LAUNCH_CHECK();
  k_partial_reduce<<<(c + 31) / 32, 1024, 0, st>>>(partial, nb, c, dgamma);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_frame_pool_fwd(const float* x, int64_t n, int32_t f, int32_t c, int32_t mode, float* out,
                                  se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0 && f >= 1 && c >= 1 && mode >= 0 && mode <= 3, "bad arguments");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(x && out, "null pointer");
  k_frame_pool_fwd<<<grid_for_elems(n * c), 256, 0, as_stream(stream)>>>(x, n, f, c, mode, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_frame_pool_bwd(const float* dout, const float* x, const float* out, int64_t n, int32_t f, int32_t c,
                                  int32_t mode, float* dx, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0 && f >= 1 && c >= 1 && mode >= 0 && mode <= 3, "bad arguments");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(dout && dx && (mode <= 1 || (x && out)), "null pointer");
  k_frame_pool_bwd<<<grid_for_elems(n * c), 256, 0, as_stream(stream)>>>(dout, x, out, n, f, c, mode, dx);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_batch_pool_fwd(const float* x, const int32_t* item_ends, int32_t n_items, int32_t c, int32_t mode, float* out,
                                  se3_stream_t stream) {
  SE3_CHECK_ARG(n_items >= 0 && c >= 1 && mode >= 0 && mode <= 3, "bad arguments");
  if (n_items == 0) return SE3_OK;
  SE3_CHECK_ARG(x && item_ends && out, "null pointer");
  k_batch_pool_fwd<<<dim3((unsigned)n_items, (unsigned)((c + 31) / 32)), 256, 0, as_stream(stream)>>>(x, item_ends, c, mode, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_batch_pool_bwd(const float* dout, const int32_t* item_ends, const int32_t* row_item, int64_t rows, int32_t c,
                                  int32_t mode, float* dx, se3_stream_t stream) {
  SE3_CHECK_ARG(rows >= 0 && c >= 1 && (mode == 0 || mode == 1 || mode == 4), "bad arguments (avg, sum or gather)");
  if (rows == 0) return SE3_OK;
  SE3_CHECK_ARG(dout && item_ends && row_item && dx, "null pointer");
  k_batch_pool_bwd<<<grid_for_elems(rows * c), 256, 0, as_stream(stream)>>>(dout, item_ends, row_item, rows, c, mode, dx);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}
