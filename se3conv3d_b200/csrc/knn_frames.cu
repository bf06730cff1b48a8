// k-NN (sweep along the widest axis, one warp per point) and reference-frame construction.
//
// Reference behaviour restated (not translated):
//   knn_query      custom_ops/knn_query/knn_query.cu:18-197  -- exact k-NN inside a batch, self
//                  included, ascending distance, first-found wins on ties (forward sweep first).
//   PCA frames     pc/RotationFunctions.py:307-406
//   SO(3) frames   pc/RotationFunctions.py:176-216, 53-82
// B200-first choices: a warp (not a thread) owns a query point, candidates are examined 32 at a
// time with coalesced float4 loads of the axis-sorted cloud, the running top-k lives in registers
// (one or two entries per lane, k <= 64) and is updated with ballot/shuffle insertions; the ordering key is
// an exact (batch, float-bits) 64-bit integer instead of the reference's rounded float key.
#include <cub/cub.cuh>
#include "common.cuh"

using namespace se3;

namespace {

struct KnnWorkspace {
  float* minmax;      // [6]
  uint64_t* keys;     // [N]
  uint64_t* keys_sorted;
  int* iota;
  int* idx_sorted;
  float4* pts_sorted;  // xyz + batch id bits
  void* cub_tmp;
  size_t cub_bytes;
};

size_t knn_cub_bytes(int64_t n) { return se3::cub_tmp_bytes(0, n); }

bool knn_layout(void* ws, size_t bytes, int64_t n, KnnWorkspace& w) {
  Arena ar(ws, bytes);
  w.minmax = ar.take<float>(8);
  w.keys = ar.take<uint64_t>(n);
  w.keys_sorted = ar.take<uint64_t>(n);
  w.iota = ar.take<int>(n);
  w.idx_sorted = ar.take<int>(n);
  w.pts_sorted = ar.take<float4>(n);
  w.cub_bytes = knn_cub_bytes(n);
  w.cub_tmp = ar.take<char>(w.cub_bytes);
  return ar.ok();
}

__device__ __forceinline__ void atomic_minf(float* a, float v) {
  // order-preserving int trick (values may be negative)
  if (v >= 0) atomicMin((int*)a, __float_as_int(v)); else atomicMax((unsigned*)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_maxf(float* a, float v) {
  if (v >= 0) atomicMax((int*)a, __float_as_int(v)); else atomicMin((unsigned*)a, __float_as_uint(v));
}

__global__ void k_minmax_init(float* mm) {
  if (threadIdx.x < 3) mm[threadIdx.x] = INFINITY;
  else if (threadIdx.x < 6) mm[threadIdx.x] = -INFINITY;
}

__global__ void k_minmax(const float* __restrict__ pts, int64_t n, float* __restrict__ mm) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = pts[3 * i + d];
      lo[d] = fminf(lo[d], v);
      hi[d] = fmaxf(hi[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomic_minf(&mm[d], lo[d]);
      atomic_maxf(&mm[3 + d], hi[d]);
    }
  }
}

__device__ __forceinline__ int sort_dim_of(const float* mm) {
  // torch::argmax(max - min): first maximal index (knn_query.cu:143-149)
  const float e0 = mm[3] - mm[0], e1 = mm[4] - mm[1], e2 = mm[5] - mm[2];
  int sd = 0;
  float best = e0;
  if (e1 > best) { best = e1; sd = 1; }
  if (e2 > best) { sd = 2; }
  return sd;
}

__global__ void k_knn_keys(const float* __restrict__ pts, const int* __restrict__ batch, int64_t n,
                           const float* __restrict__ mm, uint64_t* __restrict__ keys, int* __restrict__ iota) {
  const int sd = sort_dim_of(mm);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned u = __float_as_uint(pts[3 * i + sd]);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving float -> uint
    keys[i] = ((uint64_t)(unsigned)batch[i] << 32) | u;
    iota[i] = (int)i;
  }
}

__global__ void k_knn_gather(const float* __restrict__ pts, const int* __restrict__ batch,
                             const int* __restrict__ idx_sorted, int64_t n, float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx_sorted[i];
    out[i] = make_float4(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2], __int_as_float(batch[j]));
  }
}

__device__ __forceinline__ float comp(const float4& p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }

// Running top-k of a warp, ascending: rank r lives in slot r / 32 of lane r % 32 (KS = 1: k <= 32, KS = 2: k <= 64, the
// reference's limit, knn_query.cu:135-197).  Insertion of (d, id) = ballot for the position + a shuffle shift.
template <int KS>
struct WarpTopK {
  float best[KS];
  int idx[KS];
  float tau;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KS; ++s) { best[s] = 1e10f; idx[s] = -1; }   // knn_query.cu:38
    tau = 1e10f;
  }
  // all lanes call with the same (d, id); strict d < tau: the first found wins ties
  __device__ __forceinline__ void insert(float d, int id, int k, int lane) {
    if (!(d < tau)) return;
    int pos = 0;
#pragma unroll
    for (int s = 0; s < KS; ++s) pos += __popc(__ballot_sync(0xffffffffu, 32 * s + lane < k && best[s] <= d));
    const float wrap = __shfl_sync(0xffffffffu, best[0], 31);
    const int wrapi = __shfl_sync(0xffffffffu, idx[0], 31);
#pragma unroll
    for (int s = KS - 1; s >= 0; --s) {
      float up = __shfl_up_sync(0xffffffffu, best[s], 1);
      int upi = __shfl_up_sync(0xffffffffu, idx[s], 1);
      if (s > 0 && lane == 0) { up = wrap; upi = wrapi; }
      const int r = 32 * s + lane;
      if (r > pos) { best[s] = up; idx[s] = upi; }
      if (r == pos) { best[s] = d; idx[s] = id; }
    }
    const int last = k - 1;
    tau = __shfl_sync(0xffffffffu, (KS == 2 && last >= 32) ? best[KS - 1] : best[0], last & 31);
  }
};

// One warp per query (in sorted order).
template <int KS>
__global__ void __launch_bounds__(256) k_knn_sweep(const float4* __restrict__ ps, const int* __restrict__ idx_sorted,
                                                   int n, int k, const float* __restrict__ mm,
                                                   int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int sd = sort_dim_of(mm);
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < n; p += nwarps) {
    const float4 cur = ps[p];
    const int cur_b = __float_as_int(cur.w);
    WarpTopK<KS> top;
    top.init();
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
      // dir 0: q = p, p+1, ...   dir 1: q = p-1, p-2, ...
      bool stop = false;
      for (int64_t base = 0; !stop; base += 32) {
        const int64_t q = dir == 0 ? p + base + lane : p - 1 - base - lane;
        bool valid = q >= 0 && q < n;
        float4 c = cur;
        if (valid) {
          c = ps[q];
          valid = __float_as_int(c.w) == cur_b;
        }
        // the sweep ends at the first out-of-range / foreign-batch candidate (knn_query.cu:60,97)
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const unsigned first_bad = ~vmask;
        const unsigned run = first_bad ? ((first_bad & (0u - first_bad)) - 1u) : 0xffffffffu;  // lanes before 1st invalid
        valid = valid && ((run >> lane) & 1u);
        if (~vmask) stop = true;
        const float dx = c.x - cur.x, dy = c.y - cur.y, dz = c.z - cur.z;
        const float dist = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        const float sdist = comp(c, sd) - comp(cur, sd);
        const float sd2 = sdist * sdist;
        unsigned cand = __ballot_sync(0xffffffffu, valid && dist < top.tau);
        while (cand) {
          const int src = __ffs(cand) - 1;
          cand &= cand - 1;
          const float d = __shfl_sync(0xffffffffu, dist, src);
          const int64_t qi = dir == 0 ? p + base + src : p - 1 - base - src;
          top.insert(d, (int)qi, k, lane);
        }
        // stop once the axis distance alone exceeds the k-th best (knn_query.cu:84,121)
        if (__ballot_sync(0xffffffffu, valid && top.tau < sd2)) stop = true;
      }
    }
    const int self = idx_sorted[p];
#pragma unroll
    for (int s = 0; s < KS; ++s)
      if (32 * s + lane < k) out[(int64_t)self * k + 32 * s + lane] = top.idx[s] >= 0 ? idx_sorted[top.idx[s]] : -1;
  }
}

// Cross-cloud k-NN (pc/KnnNeighborhood.py:78-84, the torch_cluster.knn branch: global pooling convolutions between two
// different clouds): one warp per sample scans every source point of the sample's batch item (sources are grouped by
// batch item: src_ends[b] = inclusive end of item b).  out [M, k]: source ids by ascending distance, -1 padded.
template <int KS>
__global__ void __launch_bounds__(256) k_knn_cross(const float* __restrict__ src, const int* __restrict__ src_ends,
                                                   const float* __restrict__ dst, const int* __restrict__ batch_dst, int m,
                                                   int k, int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < m; p += nwarps) {
    const float cx = dst[3 * p], cy = dst[3 * p + 1], cz = dst[3 * p + 2];
    const int b = batch_dst[p];
    const int lo = b > 0 ? src_ends[b - 1] : 0, hi = src_ends[b];
    WarpTopK<KS> top;
    top.init();
    for (int base = lo; base < hi; base += 32) {
      const int q = base + lane;
      const bool valid = q < hi;
      float dist = 1e30f;
      if (valid) {
        const float dx = src[3 * (int64_t)q] - cx, dy = src[3 * (int64_t)q + 1] - cy, dz = src[3 * (int64_t)q + 2] - cz;
        dist = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
      }
      unsigned cand = __ballot_sync(0xffffffffu, valid && dist < top.tau);
      while (cand) {
        const int s = __ffs(cand) - 1;
        cand &= cand - 1;
        top.insert(__shfl_sync(0xffffffffu, dist, s), base + s, k, lane);
      }
    }
#pragma unroll
    for (int s = 0; s < KS; ++s)
      if (32 * s + lane < k) out[p * k + 32 * s + lane] = top.idx[s];
  }
}

// ---- symmetric 3x3 eigen-decomposition (cyclic Jacobi, double) -------------------------------
__device__ void jacobi3(double a[3][3], double v[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-18 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < 3; ++r) {  // A <- A J
          const double arp = a[r][p], arq = a[r][q];
          a[r][p] = c * arp - s * arq;
          a[r][q] = s * arp + c * arq;
        }
        for (int r = 0; r < 3; ++r) {  // A <- J^T A
          const double apr = a[p][r], aqr = a[q][r];
          a[p][r] = c * apr - s * aqr;
          a[q][r] = s * apr + c * aqr;
        }
        for (int r = 0; r < 3; ++r) {
          const double vrp = v[r][p], vrq = v[r][q];
          v[r][p] = c * vrp - s * vrq;
          v[r][q] = s * vrp + c * vrq;
        }
      }
  }
  w[0] = a[0][0];
  w[1] = a[1][1];
  w[2] = a[2][2];
}

// sel_u / sel_frames / sel_rec (optional, fused hierarchy builder): instead of all candidates, write the n_keep
// candidates a uniform variate selects (the permutation decode of k_frames_select) and their gather records
__global__ void k_pca_frames(const float* __restrict__ pts, const int* __restrict__ knn, int64_t n, int k,
                             int fixed_axis, float* __restrict__ frames, const float* __restrict__ sel_u, int n_keep,
                             float* __restrict__ sel_frames, float4* __restrict__ sel_rec) {
  const bool fixed = fixed_axis > 0;  // `not axis_fixed` makes axis 0 behave as "none" (RotationFunctions.py:323)
  const int nf = fixed ? 2 : 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double mean[3] = {0, 0, 0};
    for (int e = 0; e < k; ++e) {
      int j = knn[i * k + e];
      if (j < 0) j = (int)i;  // missing neighbours become self loops (RotationFunctions.py:314-317)
      for (int d = 0; d < 3; ++d) mean[d] += (fixed && d == fixed_axis) ? 0.0 : (double)pts[3 * (int64_t)j + d];
    }
    for (int d = 0; d < 3; ++d) mean[d] /= k;
    double a[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int e = 0; e < k; ++e) {
      int j = knn[i * k + e];
      if (j < 0) j = (int)i;
      double x[3];
      for (int d = 0; d < 3; ++d) x[d] = ((fixed && d == fixed_axis) ? 0.0 : (double)pts[3 * (int64_t)j + d]) - mean[d];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) a[r][c] += x[r] * x[c];
    }
    double v[3][3], w[3];
    jacobi3(a, v, w);
    // order: ascending eigenvalues (eigh); descending on the fixed-axis branch (torch.flip)
    int ord[3] = {0, 1, 2};
    for (int s = 0; s < 2; ++s)
      for (int t = 0; t < 2 - s; ++t) {
        const bool sw = fixed ? (w[ord[t]] < w[ord[t + 1]]) : (w[ord[t]] > w[ord[t + 1]]);
        if (sw) { const int tmp = ord[t]; ord[t] = ord[t + 1]; ord[t + 1] = tmp; }
      }
    double m[3][3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m[r][c] = v[r][ord[c]];
    const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
                       m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                       m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    if (det < 0)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) m[r][c] = -m[r][c];
    // proper sign flips of the columns, in itertools.product([1,-1],repeat=3) order
    const float sg4[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
    const float sg2[2][3] = {{1, 1, 1}, {-1, -1, 1}};
    int perm[4] = {0, 1, 2, 3};
    if (sel_frames && sel_u) {
      // one uniform variate -> a permutation of the nf candidates (Fisher-Yates), as k_frames_select
      int total = 1;
      for (int q = 2; q <= nf; ++q) total *= q;
      int code = min((int)(sel_u[i] * (float)total), total - 1);
      for (int q = 0; q < nf - 1; ++q) {
        const int span = nf - q;
        const int pick = q + code % span;
        code /= span;
        const int tmp = perm[q]; perm[q] = perm[pick]; perm[pick] = tmp;
      }
    }
    const int n_out = sel_frames ? n_keep : nf;
    for (int fo = 0; fo < n_out; ++fo) {
      const int f = sel_frames ? perm[fo] : fo;
      float o[3][3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o[r][c] = (float)m[r][c] * (fixed ? sg2[f][c] : sg4[f][c]);
      float fr[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          int cc = c;
          if (fixed && fixed_axis == 1) cc = (c == 0) ? 0 : (c == 1 ? 2 : 1);  // columns [0,2,1]
          float val = o[r][cc];
          if (fixed && fabsf(val) < 1e-6f) val = 0.0f;
          fr[r * 3 + c] = val;
        }
      float* dst = sel_frames ? sel_frames + (i * n_keep + fo) * 9 : frames + (i * nf + fo) * 9;
#pragma unroll
      for (int q = 0; q < 9; ++q) dst[q] = fr[q];
      if (sel_frames) {
        float4* rc = sel_rec + 3 * (i * n_keep + fo);
        rc[0] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], fr[0]);
        rc[1] = make_float4(fr[1], fr[2], fr[3], fr[4]);
        rc[2] = make_float4(fr[5], fr[6], fr[7], fr[8]);
      }
    }
  }
}

__global__ void k_quat_frames(const float* __restrict__ q, int64_t n, float* __restrict__ frames) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    float o0 = q[4 * t], o1 = q[4 * t + 1], o2 = q[4 * t + 2], o3 = q[4 * t + 3];
    // random_quaternions: o / copysign(sqrt(sum o^2), o[0])   (RotationFunctions.py:194-197)
    const float s = o0 * o0 + o1 * o1 + o2 * o2 + o3 * o3;
    float nrm = sqrtf(s);
    if ((nrm < 0) != (o0 < 0)) nrm = -nrm;
    const float r = o0 / nrm, i = o1 / nrm, j = o2 / nrm, k = o3 / nrm;
    // quaternion_to_matrix (RotationFunctions.py:53-82)
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    float* o = frames + 9 * t;
    o[0] = 1 - two_s * (j * j + k * k);
    o[1] = two_s * (i * j - k * r);
    o[2] = two_s * (i * k + j * r);
    o[3] = two_s * (i * j + k * r);
    o[4] = 1 - two_s * (i * i + k * k);
    o[5] = two_s * (j * k - i * r);
    o[6] = two_s * (i * k - j * r);
    o[7] = two_s * (j * k + i * r);
    o[8] = 1 - two_s * (i * i + j * j);
  }
}

inline int grid_for(int64_t n, int block) {
  int64_t blocks = (n + block - 1) / block;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" size_t se3_knn_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  Arena ar(nullptr, 0);
  ar.take<float>(8);
  ar.take<uint64_t>(n);
  ar.take<uint64_t>(n);
  ar.take<int>(n);
  ar.take<int>(n);
  ar.take<float4>(n);
  ar.take<char>(knn_cub_bytes(n));
  return ar.off + 256;
}

extern "C" int se3_knn_query(const float* pts, const int32_t* batch_ids, int64_t n, int32_t k, void* workspace,
                             size_t workspace_bytes, int32_t* out, se3_stream_t stream) {
  return se3::knn_query_impl(pts, batch_ids, n, k, workspace, workspace_bytes, out, 0, 0, stream);
}

int se3::knn_query_impl(const float* pts, const int32_t* batch_ids, int64_t n, int32_t k, void* workspace,
                        size_t workspace_bytes, int32_t* out, int32_t n_batches, int32_t max_seg, se3_stream_t stream,
                        const float* raw_min, const float* raw_max) {
  SE3_CHECK_ARG(n >= 0 && n < (1ll << 31), "bad n");
  SE3_CHECK_ARG(k >= 1 && k <= 64, "k must be in 1..64");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(pts && batch_ids && workspace && out, "null pointer");
  cudaStream_t st = as_stream(stream);
  KnnWorkspace w;
  if (!knn_layout(workspace, workspace_bytes, n, w)) {
    set_error("se3_knn_query: workspace too small");
    return SE3_EWORKSPACE;
  }
  if (raw_min && raw_max && seg_build_possible(n_batches, max_seg)) {
    // per-item boxes are known and every batch item fits a CTA: sweep axis, keys, sort and gather in one launch
    if (int rc = knn_sorted_fused(pts, batch_ids, n, raw_min, raw_max, w.idx_sorted, w.pts_sorted, w.minmax, n_batches,
                                  max_seg, stream))
      return rc;
    if (k <= 32) k_knn_sweep<1><<<grid_for(n * 32, 256), 256, 0, st>>>(w.pts_sorted, w.idx_sorted, (int)n, k, w.minmax, out);
    else k_knn_sweep<2><<<grid_for(n * 32, 256), 256, 0, st>>>(w.pts_sorted, w.idx_sorted, (int)n, k, w.minmax, out);
    SE3_LAUNCH_CHECK();
    return SE3_OK;
  }
  k_minmax_init<<<1, 32, 0, st>>>(w.minmax);
  SE3_LAUNCH_CHECK();
  k_minmax<<<grid_for(n, 256), 256, 0, st>>>(pts, n, w.minmax);
  SE3_LAUNCH_CHECK();
  k_knn_keys<<<grid_for(n, 256), 256, 0, st>>>(pts, batch_ids, n, w.minmax, w.keys, w.iota);
  SE3_LAUNCH_CHECK();
  // per batch item the keys differ in the low 32 bits only (the float bits of the sweep coordinate)
  if (int rc = sort_keys_u64(w.keys, w.iota, batch_ids, n, n_batches, max_seg, w.keys_sorted, w.idx_sorted,
                             (max_seg > 0 && max_seg <= kSegSortMax) ? 32 : 64, w.cub_tmp, w.cub_bytes, st))
    return rc;
  k_knn_gather<<<grid_for(n, 256), 256, 0, st>>>(pts, batch_ids, w.idx_sorted, n, w.pts_sorted);
  SE3_LAUNCH_CHECK();
  if (k <= 32) k_knn_sweep<1><<<grid_for(n * 32, 256), 256, 0, st>>>(w.pts_sorted, w.idx_sorted, (int)n, k, w.minmax, out);
  else k_knn_sweep<2><<<grid_for(n * 32, 256), 256, 0, st>>>(w.pts_sorted, w.idx_sorted, (int)n, k, w.minmax, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_knn_cross(const float* pts_src, const int32_t* src_batch_ends, const float* pts_dst, const int32_t* batch_dst,
                             int64_t m, int32_t k, int32_t* out, se3_stream_t stream) {
  SE3_CHECK_ARG(m >= 0 && m < ((int64_t)1 << 31) && k >= 1 && k <= 64, "bad sizes (k must be in 1..64)");
  if (m == 0) return SE3_OK;
  SE3_CHECK_ARG(pts_src && src_batch_ends && pts_dst && batch_dst && out, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (k <= 32) k_knn_cross<1><<<grid_for(m * 32, 256), 256, 0, st>>>(pts_src, src_batch_ends, pts_dst, batch_dst, (int)m, k, out);
  else k_knn_cross<2><<<grid_for(m * 32, 256), 256, 0, st>>>(pts_src, src_batch_ends, pts_dst, batch_dst, (int)m, k, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_pca_frames(const float* pts, const int32_t* knn, int64_t n, int32_t k, int32_t fixed_axis,
                              float* frames_out, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0 && k >= 1 && fixed_axis >= -1 && fixed_axis <= 2, "bad arguments");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(pts && knn && frames_out, "null pointer");
  k_pca_frames<<<grid_for(n, 128), 128, 0, as_stream(stream)>>>(pts, knn, n, k, fixed_axis, frames_out, nullptr, 0, nullptr,
                                                                nullptr);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// PCA frames + candidate selection + gather records in one launch (fused hierarchy builder)
int se3::pca_frames_select_pack(const float* pts, const int32_t* knn, int64_t n, int32_t k, int32_t fixed_axis,
                                const float* u, int32_t n_keep, float* frames_out, float* rec_out, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0 && k >= 1 && fixed_axis >= -1 && fixed_axis <= 2 && n_keep >= 1 && n_keep <= (fixed_axis > 0 ? 2 : 4),
                "bad arguments");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(pts && knn && frames_out && rec_out && (reinterpret_cast<uintptr_t>(rec_out) & 15) == 0, "bad pointer");
  k_pca_frames<<<grid_for(n, 128), 128, 0, as_stream(stream)>>>(pts, knn, n, k, fixed_axis, nullptr, u, n_keep, frames_out,
                                                                reinterpret_cast<float4*>(rec_out));
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_quat_frames(const float* q, int64_t n, float* frames_out, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0, "bad n");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(q && frames_out, "null pointer");
  k_quat_frames<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(q, n, frames_out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}
