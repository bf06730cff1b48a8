// Grid keys, grid-hashed ball query -> CSR, CSR transpose, segment pooling.
//
// Replaces custom_ops/ball_query/*.cu of the reference (see include/se3conv3d_b200.h for the
// per-entry-point citations).  Design differences (B200-first, not a translation):
//   * no per-(batch,x,y) "tube" table (build_grid_ds.cu) and therefore no host knowledge of the
//     grid extents: tube ranges are found by two binary searches over the sorted 64-bit keys,
//     so the whole query runs without a single host synchronisation;
//   * one warp per sample scans its <=9 key ranges with coalesced float4 loads of the
//     key-sorted source points; hits are counted / compacted with ballots, no atomics, so the
//     neighbour order inside a row is deterministic;
//   * the count pass writes the inclusive row ends through a device-wide scan (CUB), the fill
//     pass reuses the cached ranges.
// All float predicates follow the reference operation order exactly (bit-exact indices):
//   cell  : floorf((p - min) * (1/cell))                 grid_utils.cuh:64-66
//   hit   : sqrtf(fma(dz,dz,fma(dy,dy,dx*dx))) < 1.0f     count_neighbors.cu:86, math_helper.cuh:304-320
//           with d = (sample - point) * (1/radius)
#include <cub/cub.cuh>
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace se3 {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
bool pdl_enabled() {
  static const bool on = !(getenv("SE3_PDL") && getenv("SE3_PDL")[0] == '0');
  return on;
}

// ---- optional per-kernel device timing of the dominant kernels (bench.py's roofline): CUDA events on the
// launching stream around each launch, summed on read.  Off by default: no events, no overhead.
static bool g_prof_on = false;
struct ProfRec { int id; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
bool profile_enabled() { return g_prof_on; }
void profile_begin(int id, cudaStream_t st, void** handle) {
  *handle = nullptr;
  if (!g_prof_on) return;
  ProfRec* r = new ProfRec;
  r->id = id;
  if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) { delete r; return; }
  cudaEventRecord(r->a, st);
  *handle = r;
}
void profile_end(void* handle, cudaStream_t st) {
  if (!handle) return;
  ProfRec* r = reinterpret_cast<ProfRec*>(handle);
  cudaEventRecord(r->b, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(*r);
  delete r;
}
}  // namespace se3

extern "C" void se3_profile_enable(int32_t on) { se3::g_prof_on = on != 0; }
// ms_out / count_out [SE3_PROF_KERNELS]: accumulated device time and launches per profiled kernel since the last
// read (synchronises the device, then resets).
extern "C" int se3_profile_read(double* ms_out, int64_t* count_out) {
  using namespace se3;
  for (int i = 0; i < SE3_PROF_KERNELS; ++i) { ms_out[i] = 0.0; count_out[i] = 0; }
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("se3_profile_read: device synchronise failed"); return SE3_ECUDA; }
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.id >= 0 && r.id < SE3_PROF_KERNELS) {
      ms_out[r.id] += ms;
      count_out[r.id] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return SE3_OK;
}

using namespace se3;

// CUB temporary-storage sizes, memoised per (kind, next power of two >= n): the size queries walk the device
// attributes / kernel occupancy and cost microseconds each -- the builders ask dozens of times per call.
//   kind 0: SortPairs(64-bit key, int value)   1: InclusiveSum(int)   2: SortPairs(int key, int value)
size_t se3::cub_tmp_bytes(int kind, int64_t n) {
  static size_t memo[3][40] = {};
  int lg = 10;
  while ((1ll << lg) < n && lg < 31) ++lg;
  size_t& m = memo[kind][lg];
  if (m == 0) {
    const int nn = (int)((1ll << lg) > 2147483647ll ? 2147483647ll : (1ll << lg));
    size_t b = 0;
    if (kind == 0)
      cub::DeviceRadixSort::SortPairs(nullptr, b, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr, nn);
    else if (kind == 1)
      cub::DeviceScan::InclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, nn);
    else
      cub::DeviceRadixSort::SortPairs(nullptr, b, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, nn);
    m = b ? b : 1;
  }
  return m;
}

extern "C" int se3_abi_version(void) { return 1; }
extern "C" const char* se3_last_error(void) { return se3::g_err; }
extern "C" int64_t se3_launch_count(void) { return se3::g_launches.load(); }

// ---------------------------------------------------------------------------------------------
// keys
// ---------------------------------------------------------------------------------------------
struct GridParams {
  int nx, ny, nz;
  float ix, iy, iz;  // reciprocal cell size
};

__device__ __forceinline__ GridParams load_grid(const int* __restrict__ num_cells,
                                                const float* __restrict__ cell_size) {
  GridParams g;
  g.nx = num_cells[0];
  g.ny = num_cells[1];
  g.nz = num_cells[2];
  // torch::reciprocal == IEEE 1.0f / x (compute_keys.cu:112)
  g.ix = __fdiv_rn(1.0f, cell_size[0]);
  g.iy = __fdiv_rn(1.0f, cell_size[1]);
  g.iz = __fdiv_rn(1.0f, cell_size[2]);
  return g;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ void point_cell(const GridParams& g, float px, float py, float pz,
                                           float mx, float my, float mz, int& cx, int& cy, int& cz) {
  // (p - min) * inv : a subtraction followed by a multiplication, nothing to contract
  cx = clampi((int)floorf(__fmul_rn(__fsub_rn(px, mx), g.ix)), 0, g.nx - 1);
  cy = clampi((int)floorf(__fmul_rn(__fsub_rn(py, my), g.iy)), 0, g.ny - 1);
  cz = clampi((int)floorf(__fmul_rn(__fsub_rn(pz, mz), g.iz)), 0, g.nz - 1);
}

__device__ __forceinline__ int64_t cell_key(const GridParams& g, int b, int cx, int cy, int cz) {
  return (((int64_t)b * g.nx + cx) * g.ny + cy) * g.nz + cz;
}

__global__ void k_compute_keys(const float* __restrict__ pts, const int* __restrict__ batch, int64_t n,
                               const float* __restrict__ aabb_min, const int* __restrict__ num_cells,
                               const float* __restrict__ cell_size, int64_t* __restrict__ keys,
                               int* __restrict__ iota) {
  const GridParams g = load_grid(num_cells, cell_size);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = batch[i];
    int cx, cy, cz;
    point_cell(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], aabb_min[3 * b], aabb_min[3 * b + 1],
               aabb_min[3 * b + 2], cx, cy, cz);
    keys[i] = cell_key(g, b, cx, cy, cz);
    if (iota) iota[i] = (int)i;
  }
}

static inline int grid_for(int64_t n, int block) {
  int64_t blocks = (n + block - 1) / block;
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

extern "C" int se3_compute_keys(const float* pts, const int32_t* batch_ids, int64_t n, const float* aabb_min,
                                const int32_t* num_cells, const float* cell_size, int64_t* keys_out,
                                se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(pts && batch_ids && aabb_min && num_cells && cell_size && keys_out, "null pointer");
  k_compute_keys<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(pts, batch_ids, n, aabb_min, num_cells,
                                                                 cell_size, keys_out, nullptr);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// ---------------------------------------------------------------------------------------------
// bounding box / grid extents, dense cell ranks, frame selection
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_f32(float* a, float v) {
  if (v >= 0) atomicMin((int*)a, __float_as_int(v)); else atomicMax((unsigned*)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float* a, float v) {
  if (v >= 0) atomicMax((int*)a, __float_as_int(v)); else atomicMin((unsigned*)a, __float_as_uint(v));
}

__global__ void k_bbox_init(float* mn, float* mx, int nb3) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb3; i += gridDim.x * blockDim.x) {
    mn[i] = INFINITY;
    mx[i] = -INFINITY;
  }
}

__global__ void k_bbox_reduce(const float* __restrict__ pts, const int* __restrict__ batch, int64_t n, float* mn,
                              float* mx) {
  const int lane = threadIdx.x & 31;
  for (int64_t base = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) - lane; base < n;
       base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + lane;
    const bool ok = i < n;
    const int b = ok ? batch[i] : -1;
    float v[3] = {0.f, 0.f, 0.f};
    if (ok) { v[0] = pts[3 * i]; v[1] = pts[3 * i + 1]; v[2] = pts[3 * i + 2]; }
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool uniform = __all_sync(0xffffffffu, b == b0);
    if (uniform) {  // the common case: batch ids are contiguous
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float lo = v[d], hi = v[d];
        for (int o = 16; o > 0; o >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
          hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { atomic_min_f32(mn + 3 * b0 + d, lo); atomic_max_f32(mx + 3 * b0 + d, hi); }
      }
    } else if (ok) {
#pragma unroll
      for (int d = 0; d < 3; ++d) { atomic_min_f32(mn + 3 * b + d, v[d]); atomic_max_f32(mx + 3 * b + d, v[d]); }
    }
  }
}

// one CTA per batch item (batch ids are non-decreasing): no initialisation pass, no atomics
__global__ void __launch_bounds__(256) k_bbox_seg(const float* __restrict__ pts, const int* __restrict__ batch, int n,
                                                  float* __restrict__ mn, float* __restrict__ mx) {
  __shared__ int s_lo, s_hi;
  __shared__ float s_red[8][6];
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (batch[mid] < b) lo = mid + 1; else hi = mid; }
    s_lo = lo;
    hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (batch[mid] < b + 1) lo = mid + 1; else hi = mid; }
    s_hi = lo;
  }
  __syncthreads();
  float v[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int i = s_lo + threadIdx.x; i < s_hi; i += blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float x = pts[3 * (int64_t)i + d];
      v[d] = fminf(v[d], x);
      v[3 + d] = fmaxf(v[3 + d], x);
    }
  }
#pragma unroll
  for (int d = 0; d < 6; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      const float y = __shfl_xor_sync(0xffffffffu, v[d], o);
      v[d] = d < 3 ? fminf(v[d], y) : fmaxf(v[d], y);
    }
  if ((threadIdx.x & 31) == 0)
    for (int d = 0; d < 6; ++d) s_red[threadIdx.x >> 5][d] = v[d];
  __syncthreads();
  if (threadIdx.x < 6) {
    const int d = threadIdx.x;
    float r = s_red[0][d];
    for (int w = 1; w < 8; ++w) r = d < 3 ? fminf(r, s_red[w][d]) : fmaxf(r, s_red[w][d]);
    if (d < 3) mn[3 * b + d] = r; else mx[3 * b + d - 3] = r;
  }
}

// raw per-batch bounding box (no padding): mn/mx [B,3]; batches without points keep (+inf, -inf)
extern "C" int se3_bbox(const float* pts, const int32_t* batch_ids, int64_t n, int32_t n_batches, float* min_out,
                        float* max_out, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0 && n_batches >= 1 && pts && batch_ids && min_out && max_out, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (n < (1ll << 31) && n_batches <= 65535 && n / n_batches <= 16384) {
    k_bbox_seg<<<n_batches, 256, 0, st>>>(pts, batch_ids, (int)n, min_out, max_out);
    SE3_LAUNCH_CHECK();
    return SE3_OK;
  }
  k_bbox_init<<<(n_batches * 3 + 127) / 128, 128, 0, st>>>(min_out, max_out, n_batches * 3);
  SE3_LAUNCH_CHECK();
  if (n > 0) {
    k_bbox_reduce<<<grid_for(n, 256), 256, 0, st>>>(pts, batch_ids, n, min_out, max_out);
    SE3_LAUNCH_CHECK();
  }
  return SE3_OK;
}

__global__ void k_bbox_finalize(const float* raw_mn, const float* raw_mx, float* mn, float* mx,
                                int nb, float cell, float max_pad, int* num_cells) {
  // one warp; raw box in (may alias the padded box out); batches without points keep (+inf, -inf) in the raw box,
  // get a zero box and do not contribute
  const int lane = threadIdx.x;
  const float inv = __fdiv_rn(1.0f, cell);
  int best[3] = {1, 1, 1};
  bool any = false;
  for (int b = lane; b < nb; b += 32) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float lo = raw_mn[3 * b + d], hi = raw_mx[3 * b + d];
      if (lo <= hi) {
        const float plo = __fsub_rn(lo, 1e-6f), phi = __fadd_rn(hi, max_pad);
        mn[3 * b + d] = plo;
        mx[3 * b + d] = phi;
        const int c = (int)__fmul_rn(__fsub_rn(phi, plo), inv) + 1;
        best[d] = any ? max(best[d], c) : c;
      } else {
        mn[3 * b + d] = 0.0f;
        mx[3 * b + d] = 0.0f;
      }
    }
    any = true;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    int v = any ? best[d] : INT_MIN;
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) num_cells[d] = v == INT_MIN ? 1 : v;
  }
}

// padded box + grid extents of a raw bounding box (se3_bbox) for one cell size: same arithmetic as se3_grid_setup
extern "C" int se3_grid_extents(const float* raw_min, const float* raw_max, int32_t n_batches, float cell,
                                float max_pad, float* min_pt_out, float* max_pt_out, int32_t* num_cells_out,
                                se3_stream_t stream) {
  SE3_CHECK_ARG(n_batches >= 1 && cell > 0.0f && raw_min && raw_max && min_pt_out && max_pt_out && num_cells_out,
                "bad arguments");
  cudaStream_t st = as_stream(stream);
  k_bbox_finalize<<<1, 32, 0, st>>>(raw_min, raw_max, min_pt_out, max_pt_out, n_batches, cell, max_pad, num_cells_out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_grid_setup(const float* pts, const int32_t* batch_ids, int64_t n, int32_t n_batches, float cell,
                              float max_pad, float* min_pt_out, float* max_pt_out, int32_t* num_cells_out,
                              se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0 && n_batches >= 1 && cell > 0.0f, "bad arguments");
  SE3_CHECK_ARG(pts && batch_ids && min_pt_out && max_pt_out && num_cells_out, "null pointer");
  cudaStream_t st = as_stream(stream);
  k_bbox_init<<<(n_batches * 3 + 127) / 128, 128, 0, st>>>(min_pt_out, max_pt_out, n_batches * 3);
  SE3_LAUNCH_CHECK();
  if (n > 0) {
    k_bbox_reduce<<<grid_for(n, 256), 256, 0, st>>>(pts, batch_ids, n, min_pt_out, max_pt_out);
    SE3_LAUNCH_CHECK();
  }
  k_bbox_finalize<<<1, 32, 0, st>>>(min_pt_out, max_pt_out, min_pt_out, max_pt_out, n_batches, cell, max_pad,
                                    num_cells_out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

__global__ void k_cell_keys(const float* __restrict__ pts, const int* __restrict__ batch, int64_t n,
                            const float* __restrict__ aabb_min, const int* __restrict__ num_cells, float cell,
                            int64_t* __restrict__ keys, int* __restrict__ iota, int* __restrict__ zero, int n_zero) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += gridDim.x * blockDim.x) zero[i] = 0;
  GridParams g;
  g.nx = num_cells[0]; g.ny = num_cells[1]; g.nz = num_cells[2];
  g.ix = g.iy = g.iz = __fdiv_rn(1.0f, cell);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = batch[i];
    int cx, cy, cz;
    point_cell(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], aabb_min[3 * b], aabb_min[3 * b + 1],
               aabb_min[3 * b + 2], cx, cy, cz);
    keys[i] = cell_key(g, b, cx, cy, cz);
    iota[i] = (int)i;
  }
}

__global__ void k_cell_flags(const int64_t* __restrict__ keys_sorted, int64_t n, int* __restrict__ flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flags[i] = (i == 0 || keys_sorted[i] != keys_sorted[i - 1]) ? 1 : 0;
}

// batch / batch_cells (optional): batch_cells[b] = number of cells of batch item b, i.e. the size of item b in the
// pooled cloud (the caller zeroes it; the host uses the largest one to pick the per-item sort of the next level)
__global__ void k_cell_ranks(const int* __restrict__ rank1, const int* __restrict__ idx_sorted, int64_t n,
                             int64_t* __restrict__ cell_ids, int64_t* __restrict__ sorted_ids,
                             int* __restrict__ cell_ends, int64_t* __restrict__ m_out,
                             const int* __restrict__ batch, int* __restrict__ batch_cells) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = rank1[i] - 1;
    const int src = idx_sorted[i];
    cell_ids[src] = r;
    sorted_ids[i] = src;
    if (i == n - 1 || rank1[i + 1] - 1 != r) cell_ends[r] = (int)(i + 1);
    if (i == n - 1) *m_out = r + 1;
    if (batch_cells) {
      // first / last sorted position of a batch item: +-(rank) contributions give its cell count without ordering
      const int b = batch[src];
      if (i == n - 1 || batch[idx_sorted[i + 1]] != b) atomicAdd(batch_cells + b, r + 1);
      if (i == 0 || batch[idx_sorted[i - 1]] != b) atomicAdd(batch_cells + b, -r);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-batch-item sorted structures
// ---------------------------------------------------------------------------------------------
// Batch ids are non-decreasing and every key carries its batch id in the leading position, so sorting each batch
// item's slice on its own and leaving the slices in place IS the global stable sort.  A cloud of a few thousand
// points per item (DFAUST: 6890, pooled levels: hundreds) then needs ONE launch (k_seg_build below) instead of the
// histogram + scan + one pass per 8 key bits of the device-wide sort, whose launches -- not its bandwidth --
// dominate at this size.  Values are the global positions (the sort produces the permutation).
__device__ __forceinline__ int lower_bound_i32(const int* __restrict__ a, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
// one launch per sorted structure (fused hierarchy builder, clouds whose batch items fit a CTA)
// ---------------------------------------------------------------------------------------------
// CTA b builds batch item b's slice of a sorted structure from the raw per-item boxes: grid extents (KIND 0) or
// the sweep axis (KIND 1), the keys, the per-item sort and the key-ordered coordinates -- the work of se3_grid_extents
// / k_minmax, k_cell_keys / k_compute_keys / k_knn_keys, the radix sort and the gather kernels in ONE launch.  The
// arithmetic is the same as in those kernels (same rounding, same order), so the results are bit-identical.
//   KIND 0: voxel keys of a grid with cell size `cell` (pooling grids and ball-query sources)
//   KIND 1: kNN sweep order (batch, coordinate along the widest axis of the whole cloud)
struct SegBuildArgs {
  const float* pts;
  const int* batch;
  int n, n_batches;
  const float* raw_min;  // [B,3] raw boxes (se3_bbox)
  const float* raw_max;
  float cell, max_pad;   // KIND 0
  uint64_t* keys_sorted;  // KIND 0: [n] full keys, ascending
  int* idx_sorted;        // [n] sorted position -> point
  float4* pts_sorted;     // [n] xyz (+ batch id bits, KIND 1) in sorted order; optional for KIND 0
  float* mn_out;          // KIND 0: padded boxes [B,3] and grid extents [3] (written by CTA 0)
  float* mx_out;
  int* nc_out;
  float* mm_out;          // KIND 1: min / max of the whole cloud [6] (written by CTA 0)
  int* zero;              // optional: n_zero ints cleared by CTA 0
  int n_zero;
};

template <int THREADS, int ITEMS, int KIND>
__global__ void __launch_bounds__(THREADS) k_seg_build(const SegBuildArgs a) {
  using Sort = cub::BlockRadixSort<uint64_t, THREADS, ITEMS, int>;
  extern __shared__ __align__(16) unsigned char seg_smem[];
  typename Sort::TempStorage& tmp = *reinterpret_cast<typename Sort::TempStorage*>(seg_smem);
  __shared__ int s_lo, s_hi, s_nc[3];
  __shared__ float s_mm[6];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    s_lo = lower_bound_i32(a.batch, a.n, b);
    s_hi = lower_bound_i32(a.batch, a.n, b + 1);
    s_nc[0] = s_nc[1] = s_nc[2] = 1;
    s_mm[0] = s_mm[1] = s_mm[2] = INFINITY;
    s_mm[3] = s_mm[4] = s_mm[5] = -INFINITY;
  }
  if (b == 0)
    for (int i = tid; i < a.n_zero; i += THREADS) a.zero[i] = 0;
  __syncthreads();
  const float inv = __fdiv_rn(1.0f, a.cell);
  // ---- extents over all batch items (every CTA needs them; CTA 0 also publishes them)
  for (int bb = tid; bb < a.n_batches; bb += THREADS) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float lo = a.raw_min[3 * bb + d], hi = a.raw_max[3 * bb + d];
      if (KIND == 0) {
        float plo = 0.0f, phi = 0.0f;
        if (lo <= hi) {
          plo = __fsub_rn(lo, 1e-6f);
          phi = __fadd_rn(hi, a.max_pad);
          atomicMax(&s_nc[d], (int)__fmul_rn(__fsub_rn(phi, plo), inv) + 1);
        }
        if (b == 0) {
          a.mn_out[3 * bb + d] = plo;
          a.mx_out[3 * bb + d] = phi;
        }
      } else if (lo <= hi) {
        atomic_min_f32(&s_mm[d], lo);
        atomic_max_f32(&s_mm[3 + d], hi);
      }
    }
  }
  __syncthreads();
  if (b == 0 && tid < 3 && KIND == 0) a.nc_out[tid] = s_nc[tid];
  if (b == 0 && tid < 6 && KIND == 1) a.mm_out[tid] = s_mm[tid];
  const int lo = s_lo, cnt = s_hi - lo;
  if (cnt <= 0) return;
  uint64_t k[ITEMS];
  int v[ITEMS];
  int end_bit = 32;
  uint64_t key_base = 0;
  if (KIND == 0) {
    GridParams g;
    g.nx = s_nc[0]; g.ny = s_nc[1]; g.nz = s_nc[2];
    g.ix = g.iy = g.iz = inv;
    float mn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float rl = a.raw_min[3 * b + d], rh = a.raw_max[3 * b + d];
      mn[d] = rl <= rh ? __fsub_rn(rl, 1e-6f) : 0.0f;
    }
    const uint64_t cells = (uint64_t)g.nx * (uint64_t)g.ny * (uint64_t)g.nz;
    key_base = (uint64_t)b * cells;
    end_bit = cells > 1 ? 64 - __clzll((long long)(cells - 1)) : 1;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int p = tid * ITEMS + i;
      k[i] = ~0ull;
      v[i] = lo + p;
      if (p < cnt) {
        const int64_t q = lo + p;
        int cx, cy, cz;
        point_cell(g, a.pts[3 * q], a.pts[3 * q + 1], a.pts[3 * q + 2], mn[0], mn[1], mn[2], cx, cy, cz);
        k[i] = (uint64_t)(((int64_t)cx * g.ny + cy) * g.nz + cz);
      }
    }
    // the padding keys must stay above every real key inside the sorted bit range
    if (end_bit < 64) {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i)
        if (tid * ITEMS + i >= cnt) k[i] = (1ull << end_bit) - 1;
    }
  } else {
    // torch::argmax(max - min): first maximal index (knn_query.cu:143-149), as sort_dim_of in knn_frames.cu
    const float e0 = s_mm[3] - s_mm[0], e1 = s_mm[4] - s_mm[1], e2 = s_mm[5] - s_mm[2];
    int sd = 0;
    float best = e0;
    if (e1 > best) { best = e1; sd = 1; }
    if (e2 > best) sd = 2;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int p = tid * ITEMS + i;
      k[i] = 0xffffffffull;
      v[i] = lo + p;
      if (p < cnt) {
        unsigned u = __float_as_uint(a.pts[3 * (int64_t)(lo + p) + sd]);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving float -> uint
        k[i] = u;
      }
    }
  }
  Sort(tmp).Sort(k, v, 0, end_bit);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int p = tid * ITEMS + i;
    if (p < cnt) {
      const int src = v[i];
      if (KIND == 0) a.keys_sorted[lo + p] = key_base + k[i];
      a.idx_sorted[lo + p] = src;
      if (a.pts_sorted)
        a.pts_sorted[lo + p] = make_float4(a.pts[3 * (int64_t)src], a.pts[3 * (int64_t)src + 1], a.pts[3 * (int64_t)src + 2],
                                           KIND == 1 ? __int_as_float(b) : 0.0f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// one launch per pooling level (fused hierarchy builder, batch items that fit a CTA)
// ---------------------------------------------------------------------------------------------
// CTA b does for batch item b everything se3_grid_extents + se3_grid_cells + se3_segment_pool_f32 +
// se3_segment_first_i32 + se3_bbox do for a level: grid extents from the raw boxes, voxel keys, the per-item sort,
// the dense cell ranks (a block scan; the rank offset of item b is the number of cells of the items before it, read
// from a small look-back array the CTAs publish into -- CTAs are dispatched in index order, so the ones a CTA waits
// for are running or done), the pooled cloud (cell means, summed in sorted order like k_segment_pool) and the raw
// boxes of the pooled cloud for the next level.  Same arithmetic and order as those kernels: bit-identical results.
struct GridLevelArgs {
  const float* pts;
  const int* batch;
  int n, n_batches;
  const float* raw_min;  // [B,3] raw boxes of the cloud being pooled
  const float* raw_max;
  float cell;
  float* mn_out;  // padded boxes [B,3] + grid extents [3] (CTA 0)
  float* mx_out;
  int* nc_out;
  int64_t* cell_ids;    // [n]
  int64_t* sorted_ids;  // [n]
  int* cell_ends;       // [<= n]
  int64_t* m_out;       // number of cells = points of the pooled cloud
  int* batch_cells;     // [B] cells per batch item (optional)
  int* state;           // [B] look-back: (cells of item b) << 1 | 1, zeroed by the caller
  float* pool_pts;      // [<= n, 3]
  int* pool_batch;      // [<= n]
  float* pool_min;      // [B,3] raw boxes of the pooled cloud
  float* pool_max;
};

template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_grid_level(const GridLevelArgs a) {
  using Sort = cub::BlockRadixSort<uint64_t, THREADS, ITEMS, int>;
  using Disc = cub::BlockDiscontinuity<uint64_t, THREADS>;
  using Scan = cub::BlockScan<int, THREADS>;
  constexpr int CAP = THREADS * ITEMS;
  extern __shared__ __align__(16) unsigned char seg_smem[];
  typename Sort::TempStorage& tmp = *reinterpret_cast<typename Sort::TempStorage*>(seg_smem);
  int* s_idx = reinterpret_cast<int*>(seg_smem);  // after the sort the same bytes hold three int arrays [CAP]
  int* s_rank = s_idx + CAP;
  int* s_cend = s_rank + CAP;
  __shared__ typename Disc::TempStorage disc_tmp;
  __shared__ typename Scan::TempStorage scan_tmp;
  __shared__ int s_lo, s_hi, s_nc[3], s_off;
  __shared__ float s_red[THREADS / 32][6];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    s_lo = lower_bound_i32(a.batch, a.n, b);
    s_hi = lower_bound_i32(a.batch, a.n, b + 1);
    s_nc[0] = s_nc[1] = s_nc[2] = 1;
  }
  __syncthreads();
  const float inv = __fdiv_rn(1.0f, a.cell);
  for (int bb = tid; bb < a.n_batches; bb += THREADS) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float lo = a.raw_min[3 * bb + d], hi = a.raw_max[3 * bb + d];
      float plo = 0.0f, phi = 0.0f;
      if (lo <= hi) {
        plo = __fsub_rn(lo, 1e-6f);
        phi = __fadd_rn(hi, 1e-6f);
        atomicMax(&s_nc[d], (int)__fmul_rn(__fsub_rn(phi, plo), inv) + 1);
      }
      if (b == 0) {
        a.mn_out[3 * bb + d] = plo;
        a.mx_out[3 * bb + d] = phi;
      }
    }
  }
  __syncthreads();
  if (b == 0 && tid < 3) a.nc_out[tid] = s_nc[tid];
  const int lo = s_lo, cnt = s_hi - lo;
  int v[ITEMS], r1[ITEMS];
  int m_b = 0;
  if (cnt > 0) {
    uint64_t k[ITEMS];
    GridParams g;
    g.nx = s_nc[0]; g.ny = s_nc[1]; g.nz = s_nc[2];
    g.ix = g.iy = g.iz = inv;
    float mn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float rl = a.raw_min[3 * b + d], rh = a.raw_max[3 * b + d];
      mn[d] = rl <= rh ? __fsub_rn(rl, 1e-6f) : 0.0f;
    }
    const uint64_t cells = (uint64_t)g.nx * (uint64_t)g.ny * (uint64_t)g.nz;
    const int end_bit = cells > 1 ? 64 - __clzll((long long)(cells - 1)) : 1;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int p = tid * ITEMS + i;
      k[i] = ~0ull;  // padding: last in the sorted bit range (stable) and different from every real key as a whole
      v[i] = lo + p;
      if (p < cnt) {
        const int64_t q = lo + p;
        int cx, cy, cz;
        point_cell(g, a.pts[3 * q], a.pts[3 * q + 1], a.pts[3 * q + 2], mn[0], mn[1], mn[2], cx, cy, cz);
        k[i] = (uint64_t)(((int64_t)cx * g.ny + cy) * g.nz + cz);
      }
    }
    Sort(tmp).Sort(k, v, 0, end_bit);
    int head[ITEMS];
    Disc(disc_tmp).FlagHeads(head, k, cub::Inequality());
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
      if (tid * ITEMS + i >= cnt) head[i] = 0;
    Scan(scan_tmp).InclusiveSum(head, r1, m_b);
    __syncthreads();  // the sort's shared memory is free: sorted indices and ranks for the cell ends and the pooling
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int p = tid * ITEMS + i;
      if (p < cnt) {
        s_idx[p] = v[i];
        s_rank[p] = r1[i];
      }
    }
  }
  __syncthreads();
  // ---- rank offset of this item: cells of the items before it
  if (tid < 32) {
    if (tid == 0) *reinterpret_cast<volatile int*>(a.state + b) = (m_b << 1) | 1;
    int sum = 0;
    for (int base = 0; base < b; base += 32) {
      const int bb = base + lane;
      if (bb < b) {
        int val;
        do {
          val = *reinterpret_cast<volatile const int*>(a.state + bb);
        } while (!(val & 1));
        sum += val >> 1;
      }
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (tid == 0) s_off = sum;
  }
  __syncthreads();
  const int off = s_off;
  if (tid == 0) {
    if (a.batch_cells) a.batch_cells[b] = m_b;
    if (b == a.n_batches - 1) *a.m_out = (int64_t)off + m_b;
  }
  float bx[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
  if (cnt > 0) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int p = tid * ITEMS + i;
      if (p < cnt) {
        const int r = r1[i] - 1;
        a.cell_ids[v[i]] = off + r;
        a.sorted_ids[lo + p] = v[i];
        if (p == cnt - 1 || s_rank[p + 1] != r1[i]) {
          a.cell_ends[off + r] = lo + p + 1;
          s_cend[r] = p + 1;
        }
      }
    }
    __syncthreads();
    // ---- pooled cloud: mean of every cell in sorted order (k_segment_pool), batch id, raw box
    for (int c = tid; c < m_b; c += THREADS) {
      const int p0 = c > 0 ? s_cend[c - 1] : 0, p1 = s_cend[c];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float acc = 0.0f;
        for (int p = p0; p < p1; ++p) acc = acc + a.pts[3 * (int64_t)s_idx[p] + d];
        acc = acc / (float)(p1 - p0);
        a.pool_pts[3 * (int64_t)(off + c) + d] = acc;
        bx[d] = fminf(bx[d], acc);
        bx[3 + d] = fmaxf(bx[3 + d], acc);
      }
      a.pool_batch[off + c] = b;
    }
  }
#pragma unroll
  for (int d = 0; d < 6; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      const float y = __shfl_xor_sync(0xffffffffu, bx[d], o);
      bx[d] = d < 3 ? fminf(bx[d], y) : fmaxf(bx[d], y);
    }
  if (lane == 0)
    for (int d = 0; d < 6; ++d) s_red[tid >> 5][d] = bx[d];
  __syncthreads();
  if (tid < 6) {
    float r = s_red[0][tid];
    for (int w = 1; w < THREADS / 32; ++w) r = tid < 3 ? fminf(r, s_red[w][tid]) : fmaxf(r, s_red[w][tid]);
    if (tid < 3) a.pool_min[3 * b + tid] = r; else a.pool_max[3 * b + tid - 3] = r;
  }
}

template <int THREADS, int ITEMS>
static int launch_grid_level_cfg(const GridLevelArgs& a, cudaStream_t st) {
  using Sort = cub::BlockRadixSort<uint64_t, THREADS, ITEMS, int>;
  auto kern = k_grid_level<THREADS, ITEMS>;
  size_t smem = sizeof(typename Sort::TempStorage);
  if (smem < (size_t)3 * THREADS * ITEMS * sizeof(int)) smem = (size_t)3 * THREADS * ITEMS * sizeof(int);
  SE3_SMEM_ONCE(kern, smem);
  kern<<<a.n_batches, THREADS, smem, st>>>(a);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

template <int THREADS, int ITEMS, int KIND>
static int launch_seg_build_cfg(const SegBuildArgs& a, cudaStream_t st) {
  using Sort = cub::BlockRadixSort<uint64_t, THREADS, ITEMS, int>;
  auto kern = k_seg_build<THREADS, ITEMS, KIND>;
  const size_t smem = sizeof(typename Sort::TempStorage);
  SE3_SMEM_ONCE(kern, smem);
  kern<<<a.n_batches, THREADS, smem, st>>>(a);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

bool se3::seg_build_possible(int n_batches, int max_seg) {
  return max_seg > 0 && max_seg <= kSegSortMax && n_batches >= 1 && n_batches <= 65535;
}
// kind 0 / 1 as above; the caller has checked seg_build_possible
static int launch_seg_build(int kind, const SegBuildArgs& a, int max_seg, cudaStream_t st) {
#define SE3_SB_CASE(T, I)                                                                      \
  if (max_seg <= (T) * (I))                                                                    \
    return kind == 0 ? launch_seg_build_cfg<T, I, 0>(a, st) : launch_seg_build_cfg<T, I, 1>(a, st)
  SE3_SB_CASE(64, 4);
  SE3_SB_CASE(128, 8);
  SE3_SB_CASE(256, 8);
  SE3_SB_CASE(512, 8);
  SE3_SB_CASE(1024, 7);
#undef SE3_SB_CASE
  set_error("launch_seg_build: batch item too large");
  return SE3_EINVAL;
}

__global__ void k_batch_counts(const int* __restrict__ batch, int n, int n_batches, int* __restrict__ counts) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n_batches; b += gridDim.x * blockDim.x)
    counts[b] = lower_bound_i32(batch, n, b + 1) - lower_bound_i32(batch, n, b);
}

namespace se3 {
int batch_counts(const int32_t* batch_ids, int64_t n, int32_t n_batches, int32_t* counts_out, cudaStream_t st) {
  k_batch_counts<<<(n_batches + 127) / 128, 128, 0, st>>>(batch_ids, (int)n, n_batches, counts_out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}
size_t sort_pairs_tmp_bytes(int64_t n) { return cub_tmp_bytes(0, n); }
// keys_in [n] (64-bit, non-negative), iota [n] = 0..n-1 -> keys_out ascending (stable), idx_out the permutation.
// max_seg = the largest number of points in one batch item (0 = unknown -> device-wide sort).
int sort_keys_u64(const uint64_t* keys_in, const int* iota, const int* batch, int64_t n, int n_batches, int max_seg,
                  uint64_t* keys_out, int* idx_out, int end_bit, void* cub_tmp, size_t cub_bytes, cudaStream_t st) {
  if (end_bit <= 0 || end_bit > 64) end_bit = 64;
  (void)batch; (void)n_batches; (void)max_seg;  // per-item sorts live in k_seg_build (the fused builders)
  size_t cb = cub_bytes;
  SE3_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cb, keys_in, keys_out, iota, idx_out, (int)n, 0, end_bit, st));
  count_launch(1 + (end_bit + 7) / 8);
  return SE3_OK;
}
}  // namespace se3

struct CellsWorkspace {
  int64_t* keys;
  int64_t* keys_sorted;
  int* iota;
  int* idx_sorted;
  int* flags;
  int* rank1;
  void* cub_tmp;
  size_t cub_bytes;
};

static size_t cells_cub_bytes(int64_t n) {
  const size_t a = cub_tmp_bytes(0, n), b = cub_tmp_bytes(1, n);
  return a > b ? a : b;
}

static bool cells_layout(void* ws, size_t bytes, int64_t n, CellsWorkspace& w) {
  Arena ar(ws, bytes);
  w.keys = ar.take<int64_t>(n);
  w.keys_sorted = ar.take<int64_t>(n);
  w.iota = ar.take<int>(n);
  w.idx_sorted = ar.take<int>(n);
  w.flags = ar.take<int>(n);
  w.rank1 = ar.take<int>(n);
  w.cub_bytes = cells_cub_bytes(n);
  w.cub_tmp = ar.take<char>(w.cub_bytes);
  return ar.ok();
}

extern "C" size_t se3_grid_cells_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  Arena ar(nullptr, 0);
  ar.take<int64_t>(n); ar.take<int64_t>(n); ar.take<int>(n); ar.take<int>(n); ar.take<int>(n); ar.take<int>(n);
  ar.take<char>(cells_cub_bytes(n));
  return ar.off + 256;
}

extern "C" int se3_grid_cells(const float* pts, const int32_t* batch_ids, int64_t n, const float* min_pt,
                              const int32_t* num_cells, float cell, void* workspace, size_t workspace_bytes,
                              int64_t* cell_ids, int64_t* sorted_ids, int32_t* cell_ends, int64_t* m_out,
                              int32_t key_bits, se3_stream_t stream) {
  return se3::grid_cells_impl(pts, batch_ids, n, min_pt, num_cells, cell, workspace, workspace_bytes, cell_ids,
                              sorted_ids, cell_ends, m_out, key_bits, 0, 0, nullptr, stream);
}

int se3::grid_cells_impl(const float* pts, const int32_t* batch_ids, int64_t n, const float* min_pt,
                         const int32_t* num_cells, float cell, void* workspace, size_t workspace_bytes,
                         int64_t* cell_ids, int64_t* sorted_ids, int32_t* cell_ends, int64_t* m_out, int32_t key_bits,
                         int32_t n_batches, int32_t max_seg, int32_t* batch_cells, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0 && n < (1ll << 31) && cell > 0.0f, "bad arguments");
  SE3_CHECK_ARG(m_out, "null m_out");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    SE3_CUDA(cudaMemsetAsync(m_out, 0, sizeof(int64_t), st));
    return SE3_OK;
  }
  SE3_CHECK_ARG(pts && batch_ids && min_pt && num_cells && workspace && cell_ids && sorted_ids && cell_ends,
                "null pointer");
  CellsWorkspace w;
  if (!cells_layout(workspace, workspace_bytes, n, w)) {
    set_error("se3_grid_cells: workspace too small");
    return SE3_EWORKSPACE;
  }
  k_cell_keys<<<grid_for(n, 256), 256, 0, st>>>(pts, batch_ids, n, min_pt, num_cells, cell, w.keys, w.iota, batch_cells,
                                                batch_cells ? n_batches : 0);
  SE3_LAUNCH_CHECK();
  size_t cb = w.cub_bytes;
  const int end_bit = (key_bits > 0 && key_bits < 64) ? key_bits : 64;
  if (int rc = sort_keys_u64(reinterpret_cast<const uint64_t*>(w.keys), w.iota, batch_ids, n, n_batches, max_seg,
                             reinterpret_cast<uint64_t*>(w.keys_sorted), w.idx_sorted, end_bit, w.cub_tmp, w.cub_bytes, st))
    return rc;
  k_cell_flags<<<grid_for(n, 256), 256, 0, st>>>(w.keys_sorted, n, w.flags);
  SE3_LAUNCH_CHECK();
  cb = w.cub_bytes;
  SE3_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, cb, w.flags, w.rank1, (int)n, st));
  count_launch(1);
  k_cell_ranks<<<grid_for(n, 256), 256, 0, st>>>(w.rank1, w.idx_sorted, n, cell_ids, sorted_ids, cell_ends, m_out,
                                                 batch_ids, batch_cells);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// One pooling level in one launch (k_grid_level): grid of (pts, batch) with voxel `cell` from the raw boxes, and the
// pooled cloud with ITS raw boxes.  state: [n_batches] ints of scratch.  Capacity of the pooled arrays: n points.
int se3::grid_level_fused(const float* pts, const int32_t* batch_ids, int64_t n, const float* raw_min, const float* raw_max,
                          float cell, float* min_pt_out, float* max_pt_out, int32_t* num_cells_out, int64_t* cell_ids,
                          int64_t* sorted_ids, int32_t* cell_ends, int64_t* m_out, int32_t n_batches, int32_t max_seg,
                          int32_t* batch_cells, int32_t* state, float* pool_pts, int32_t* pool_batch, float* pool_min,
                          float* pool_max, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 1 && n < (1ll << 31) && cell > 0.0f && seg_build_possible(n_batches, max_seg), "bad arguments");
  cudaStream_t st = as_stream(stream);
  SE3_CUDA(cudaMemsetAsync(state, 0, (size_t)n_batches * sizeof(int32_t), st));
  GridLevelArgs a;
  a.pts = pts; a.batch = batch_ids; a.n = (int)n; a.n_batches = n_batches;
  a.raw_min = raw_min; a.raw_max = raw_max; a.cell = cell;
  a.mn_out = min_pt_out; a.mx_out = max_pt_out; a.nc_out = num_cells_out;
  a.cell_ids = cell_ids; a.sorted_ids = sorted_ids; a.cell_ends = cell_ends; a.m_out = m_out;
  a.batch_cells = batch_cells; a.state = state;
  a.pool_pts = pool_pts; a.pool_batch = pool_batch; a.pool_min = pool_min; a.pool_max = pool_max;
#define SE3_GL_CASE(T, I) \
  if (max_seg <= (T) * (I)) return launch_grid_level_cfg<T, I>(a, st)
  SE3_GL_CASE(64, 4);
  SE3_GL_CASE(128, 8);
  SE3_GL_CASE(256, 8);
  SE3_GL_CASE(512, 8);
  SE3_GL_CASE(1024, 7);
#undef SE3_GL_CASE
  set_error("grid_level_fused: batch item too large");
  return SE3_EINVAL;
}

__global__ void k_frames_select(const float* __restrict__ cand, const float* __restrict__ u, int64_t n, int n_cand,
                                int n_keep, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int perm[4] = {0, 1, 2, 3};
    if (u) {
      // decode one uniform variate into a permutation (Fisher-Yates over n_cand <= 4 items)
      int total = 1;
      for (int k = 2; k <= n_cand; ++k) total *= k;
      int code = min((int)(u[i] * (float)total), total - 1);
      for (int k = 0; k < n_cand - 1; ++k) {
        const int span = n_cand - k;
        const int pick = k + code % span;
        code /= span;
        const int tmp = perm[k]; perm[k] = perm[pick]; perm[pick] = tmp;
      }
    }
    for (int f = 0; f < n_keep; ++f) {
      const float* src = cand + (i * n_cand + perm[f]) * 9;
      float* dst = out + (i * n_keep + f) * 9;
#pragma unroll
      for (int k = 0; k < 9; ++k) dst[k] = src[k];
    }
  }
}

extern "C" int se3_frames_select(const float* cand, const float* u, int64_t n, int32_t n_cand, int32_t n_keep,
                                 float* out, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 0 && n_cand >= 1 && n_cand <= 4 && n_keep >= 1 && n_keep <= n_cand, "bad arguments");
  if (n == 0) return SE3_OK;
  SE3_CHECK_ARG(cand && out, "null pointer");
  k_frames_select<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(cand, u, n, n_cand, n_keep, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// ---------------------------------------------------------------------------------------------
// ball query
// ---------------------------------------------------------------------------------------------
struct BQWorkspace {
  int64_t* keys;         // [N]
  int64_t* keys_sorted;  // [N]
  int* iota;             // [N]
  int* idx_sorted;       // [N]  sorted position -> original source index
  float4* pts_sorted;    // [N]  xyz of the source points in key order
  int2* ranges;          // [M*9]
  int* counts;           // [M]
  void* cub_tmp;         // source part: radix-sort scratch
  size_t cub_bytes;
  void* scan_tmp;        // destination part: scan scratch (queries sharing a source may run on different streams)
  size_t scan_bytes;
};

static size_t bq_scan_bytes(int64_t n_dst) { return cub_tmp_bytes(1, n_dst); }

static size_t bq_cub_bytes(int64_t n_src, int64_t n_dst) {
  const size_t a = cub_tmp_bytes(0, n_src), b = cub_tmp_bytes(1, n_dst);
  return a > b ? a : b;
}

// source part: sorted keys / positions / coordinates (+ the CUB scratch); destination part: tube ranges, counts
static size_t bq_src_bytes(int64_t n_src, int64_t n_dst) {
  Arena ar(nullptr, 0);
  ar.take<int64_t>(n_src); ar.take<int64_t>(n_src); ar.take<int>(n_src); ar.take<int>(n_src); ar.take<float4>(n_src);
  ar.take<char>(bq_cub_bytes(n_src, n_dst));
  return ar.off;
}
static size_t bq_dst_bytes(int64_t n_dst) {
  Arena ar(nullptr, 0);
  ar.take<int2>(n_dst * 9); ar.take<int>(n_dst);
  ar.take<char>(bq_scan_bytes(n_dst));
  return ar.off;
}
static bool bq_layout2(void* ws_src, size_t src_bytes, void* ws_dst, size_t dst_bytes, int64_t n_src, int64_t n_dst,
                       int64_t n_dst_cub, BQWorkspace& w) {
  Arena as(ws_src, src_bytes);
  w.keys = as.take<int64_t>(n_src);
  w.keys_sorted = as.take<int64_t>(n_src);
  w.iota = as.take<int>(n_src);
  w.idx_sorted = as.take<int>(n_src);
  w.pts_sorted = as.take<float4>(n_src);
  w.cub_bytes = bq_cub_bytes(n_src, n_dst_cub);
  w.cub_tmp = as.take<char>(w.cub_bytes);
  Arena ad(ws_dst, dst_bytes);
  w.ranges = ad.take<int2>(n_dst * 9);
  w.counts = ad.take<int>(n_dst);
  w.scan_bytes = bq_scan_bytes(n_dst);
  w.scan_tmp = ad.take<char>(w.scan_bytes);
  return as.ok() && ad.ok();
}
static bool bq_layout(void* ws, size_t ws_bytes, int64_t n_src, int64_t n_dst, BQWorkspace& w) {
  const size_t sb = bq_src_bytes(n_src, n_dst);
  if (ws_bytes < sb) return false;
  return bq_layout2(ws, sb, reinterpret_cast<char*>(ws) + sb, ws_bytes - sb, n_src, n_dst, n_dst, w);
}

extern "C" size_t se3_ball_query_workspace_bytes(int64_t n_src, int64_t n_dst) {
  if (n_src < 1) n_src = 1;
  if (n_dst < 1) n_dst = 1;
  return bq_src_bytes(n_src, n_dst) + bq_dst_bytes(n_dst) + 256;
}
/* split workspace of the prepared variants: the source part is sized for queries of up to n_dst_max samples */
extern "C" size_t se3_ball_query_src_workspace_bytes(int64_t n_src, int64_t n_dst_max) {
  return bq_src_bytes(n_src < 1 ? 1 : n_src, n_dst_max < 1 ? 1 : n_dst_max) + 256;
}
extern "C" size_t se3_ball_query_dst_workspace_bytes(int64_t n_dst) { return bq_dst_bytes(n_dst < 1 ? 1 : n_dst) + 256; }

__global__ void k_gather_sorted_pts(const float* __restrict__ pts, const int* __restrict__ idx_sorted, int64_t n,
                                    float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = idx_sorted[i];
    out[i] = make_float4(pts[3 * (int64_t)j], pts[3 * (int64_t)j + 1], pts[3 * (int64_t)j + 2], 0.0f);
  }
}

__device__ __forceinline__ int lower_bound_i64(const int64_t* __restrict__ a, int n, int64_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ bool bq_hit(float sx, float sy, float sz, const float4 p, float irx, float iry,
                                       float irz) {
  // length((sample - point) * invRadius) < 1.0f, dot() accumulates x,y,z in order with FMA
  const float dx = __fmul_rn(__fsub_rn(sx, p.x), irx);
  const float dy = __fmul_rn(__fsub_rn(sy, p.y), iry);
  const float dz = __fmul_rn(__fsub_rn(sz, p.z), irz);
  const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
  return __fsqrt_rn(d2) < 1.0f;
}

// One warp per sample.  FILL=false: find the 9 tube ranges, cache them, count hits.
// FILL=true: rescan the cached ranges and write (sample, source) pairs.
struct BqScanArgs {
  const float* pts_dst;
  const int* batch_dst;
  int64_t n_dst;
  int n_src;
  const float* min_pt;
  const int* num_cells;
  const float* radius;
  const int64_t* keys_sorted;
  const float4* pts_sorted;
  const int* idx_sorted;
  int2* ranges;
  int* counts;
  const int* row_ends;
  int64_t* neighbors;
  int* col_src;
  int* edge_dst;
  int* t_cursor;
  int* t_edge;
};
constexpr int BQ_BATCH = kBqBatch;  // queries per batched launch (kernel parameter space)
struct BqScanBatch {
  BqScanArgs q[BQ_BATCH];
  int blk[BQ_BATCH + 1];  // first block of every query; unused entries repeat the total
};

template <bool FILL>
__device__ __forceinline__ void bq_scan_body(const BqScanArgs& a, int block, int nblocks) {
  const float* __restrict__ pts_dst = a.pts_dst;
  const int* __restrict__ batch_dst = a.batch_dst;
  const int64_t n_dst = a.n_dst;
  const int n_src = a.n_src;
  const float* __restrict__ min_pt = a.min_pt;
  const int* __restrict__ num_cells = a.num_cells;
  const float* __restrict__ radius = a.radius;
  const int64_t* __restrict__ keys_sorted = a.keys_sorted;
  const float4* __restrict__ pts_sorted = a.pts_sorted;
  const int* __restrict__ idx_sorted = a.idx_sorted;
  int2* __restrict__ ranges = a.ranges;
  int* __restrict__ counts = a.counts;
  const int* __restrict__ row_ends = a.row_ends;
  int64_t* __restrict__ neighbors = a.neighbors;
  int* __restrict__ col_src = a.col_src;
  int* __restrict__ edge_dst = a.edge_dst;
  int* __restrict__ t_cursor = a.t_cursor;
  int* __restrict__ t_edge = a.t_edge;
  // t_cursor (optional, fused hierarchy builder): the transposed CSR is built alongside.  Count pass: histogram of
  // the hit sources (t_cursor zeroed by the caller).  Fill pass: t_cursor holds the exclusive row starts and is
  // advanced per hit, so it ends as the inclusive row ends; t_edge receives the edge ids in arrival order
  // (k_t_rows_finish orders every row afterwards, so the result does not depend on the arrival order).
  const int lane = threadIdx.x & 31;
  const int64_t warp = (block * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)nblocks * blockDim.x) >> 5;
  const float irx = __fdiv_rn(1.0f, radius[0]), iry = __fdiv_rn(1.0f, radius[1]), irz = __fdiv_rn(1.0f, radius[2]);
  for (int64_t s = warp; s < n_dst; s += nwarps) {
    const float sx = pts_dst[3 * s], sy = pts_dst[3 * s + 1], sz = pts_dst[3 * s + 2];
    int2 my_range = make_int2(0, 0);
    if (!FILL) {
      if (lane < 9) {
        const GridParams g = load_grid(num_cells, radius);
        const int b = batch_dst[s];
        int cx, cy, cz;
        point_cell(g, sx, sy, sz, min_pt[3 * b], min_pt[3 * b + 1], min_pt[3 * b + 2], cx, cy, cz);
        const int ox = cx + lane / 3 - 1, oy = cy + lane % 3 - 1;
        if (ox >= 0 && ox < g.nx && oy >= 0 && oy < g.ny) {
          const int64_t klo = cell_key(g, b, ox, oy, max(cz - 1, 0));
          const int64_t khi = cell_key(g, b, ox, oy, min(cz + 1, g.nz - 1));
          my_range.x = lower_bound_i64(keys_sorted, n_src, klo);
          my_range.y = lower_bound_i64(keys_sorted, n_src, khi + 1);
        }
        ranges[s * 9 + lane] = my_range;
      }
    } else {
      if (lane < 9) my_range = ranges[s * 9 + lane];
    }
    int total = 0;
    int64_t out_base = 0;
    if (FILL) out_base = (s > 0) ? (int64_t)row_ends[s - 1] : 0;
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
      const int lo = __shfl_sync(0xffffffffu, my_range.x, r);
      const int hi = __shfl_sync(0xffffffffu, my_range.y, r);
      for (int base = lo; base < hi; base += 32) {
        const int p = base + lane;
        bool hit = false;
        if (p < hi) hit = bq_hit(sx, sy, sz, pts_sorted[p], irx, iry, irz);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (FILL && hit) {
          const int64_t slot = out_base + total + __popc(m & ((1u << lane) - 1u));
          if (neighbors) {
            neighbors[2 * slot] = s;
            neighbors[2 * slot + 1] = (int64_t)idx_sorted[p];
          } else {
            const int src = idx_sorted[p];
            col_src[slot] = src;
            edge_dst[slot] = (int)s;
            if (t_cursor) t_edge[atomicAdd(t_cursor + src, 1)] = (int)slot;
          }
        }
        if (!FILL && hit && t_cursor) atomicAdd(t_cursor + idx_sorted[p], 1);
        total += __popc(m);
      }
    }
    if (!FILL && lane == 0) counts[s] = total;
  }
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_bq_scan(const BqScanArgs a) {
  bq_scan_body<FILL>(a, (int)blockIdx.x, (int)gridDim.x);
}
// several queries in one launch (the fused builder batches the neighbourhoods that become possible together; a
// launch costs the host more than most of these queries cost the GPU): query q owns blocks [blk[q], blk[q + 1]) of
// the 1-D grid, as many as it would get on its own
template <bool FILL>
__global__ void __launch_bounds__(256) k_bq_scan_multi(const BqScanBatch b) {
  int q = 0;
  while (q + 1 < BQ_BATCH && (int)blockIdx.x >= b.blk[q + 1]) ++q;
  bq_scan_body<FILL>(b.q[q], (int)blockIdx.x - b.blk[q], b.blk[q + 1] - b.blk[q]);
}

__global__ void k_bq_total(const int* __restrict__ row_ends, int64_t n_dst, int64_t* __restrict__ total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total = n_dst > 0 ? (int64_t)row_ends[n_dst - 1] : 0;
}

static int bq_prepare(const float* pts_src, const int32_t* batch_src, int64_t n_src, const float* min_pt,
                      const int32_t* num_cells, const float* radius, BQWorkspace& w, int key_bits, cudaStream_t st,
                      int n_batches = 0, int max_seg = 0) {
  k_compute_keys<<<grid_for(n_src, 256), 256, 0, st>>>(pts_src, batch_src, n_src, min_pt, num_cells, radius, w.keys,
                                                       w.iota);
  SE3_LAUNCH_CHECK();
  const int end_bit = (key_bits > 0 && key_bits < 64) ? key_bits : 64;
  if (int rc = sort_keys_u64(reinterpret_cast<const uint64_t*>(w.keys), w.iota, batch_src, n_src, n_batches, max_seg,
                             reinterpret_cast<uint64_t*>(w.keys_sorted), w.idx_sorted, end_bit, w.cub_tmp, w.cub_bytes, st))
    return rc;
  k_gather_sorted_pts<<<grid_for(n_src, 256), 256, 0, st>>>(pts_src, w.idx_sorted, n_src, w.pts_sorted);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

static int bq_count(const float* pts_dst, const int32_t* batch_dst, int64_t n_src, int64_t n_dst, const float* min_pt,
                    const int32_t* num_cells, const float* radius, BQWorkspace& w, int32_t* row_ends_out,
                    int64_t* total_out, cudaStream_t st) {
  const int blocks = grid_for(n_dst * 32, 256);
  {
    const BqScanArgs qa{pts_dst, batch_dst, n_dst, (int)n_src, min_pt, num_cells, radius, w.keys_sorted, w.pts_sorted, w.idx_sorted, w.ranges, w.counts, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    k_bq_scan<false><<<blocks, 256, 0, st>>>(qa);
  }
  SE3_LAUNCH_CHECK();
  size_t cb = w.scan_bytes;
  SE3_CUDA(cub::DeviceScan::InclusiveSum(w.scan_tmp, cb, w.counts, row_ends_out, (int)n_dst, st));
  count_launch(1);
  k_bq_total<<<1, 32, 0, st>>>(row_ends_out, n_dst, total_out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_ball_query_count(const float* pts_src, const float* pts_dst, const int32_t* batch_src,
                                    const int32_t* batch_dst, int64_t n_src, int64_t n_dst, const float* min_pt,
                                    const int32_t* num_cells, const float* radius, void* workspace,
                                    size_t workspace_bytes, int32_t* row_ends_out, int64_t* total_out,
                                    se3_stream_t stream) {
  SE3_CHECK_ARG(n_src >= 0 && n_dst >= 0 && n_src < (1ll << 31) && n_dst < (1ll << 31), "bad sizes");
  SE3_CHECK_ARG(total_out, "null total_out");
  cudaStream_t st = as_stream(stream);
  if (n_dst == 0 || n_src == 0) {
    if (n_dst > 0) SE3_CUDA(cudaMemsetAsync(row_ends_out, 0, n_dst * sizeof(int32_t), st));
    SE3_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    return SE3_OK;
  }
  SE3_CHECK_ARG(pts_src && pts_dst && batch_src && batch_dst && min_pt && num_cells && radius && row_ends_out,
                "null pointer");
  BQWorkspace w;
  if (!bq_layout(workspace, workspace_bytes, n_src, n_dst, w)) {
    set_error("se3_ball_query_count: workspace too small");
    return SE3_EWORKSPACE;
  }
  if (int rc = bq_prepare(pts_src, batch_src, n_src, min_pt, num_cells, radius, w, 0, st)) return rc;
  return bq_count(pts_dst, batch_dst, n_src, n_dst, min_pt, num_cells, radius, w, row_ends_out, total_out, st);
}

/* Prepared variants (fused hierarchy builder): one sorted source structure per (source cloud, radius) serves
 * every query against it.  key_bits bounds the radix sort (0 = all 64 bits). */
extern "C" int se3_ball_query_prepare(const float* pts_src, const int32_t* batch_src, int64_t n_src, int64_t n_dst_max,
                                      const float* min_pt, const int32_t* num_cells, const float* radius,
                                      void* ws_src, size_t ws_src_bytes, int32_t key_bits, se3_stream_t stream) {
  return se3::ball_query_prepare_impl(pts_src, batch_src, n_src, n_dst_max, min_pt, num_cells, radius, ws_src,
                                      ws_src_bytes, key_bits, 0, 0, stream);
}

int se3::ball_query_prepare_impl(const float* pts_src, const int32_t* batch_src, int64_t n_src, int64_t n_dst_max,
                                 const float* min_pt, const int32_t* num_cells, const float* radius, void* ws_src,
                                 size_t ws_src_bytes, int32_t key_bits, int32_t n_batches, int32_t max_seg,
                                 se3_stream_t stream) {
  SE3_CHECK_ARG(n_src >= 0 && n_src < (1ll << 31), "bad sizes");
  if (n_src == 0) return SE3_OK;
  SE3_CHECK_ARG(pts_src && batch_src && min_pt && num_cells && radius && ws_src, "null pointer");
  BQWorkspace w;
  char dummy[1];
  if (!bq_layout2(ws_src, ws_src_bytes, dummy, (size_t)1 << 60, n_src, 0, n_dst_max < 1 ? 1 : n_dst_max, w)) {
    set_error("se3_ball_query_prepare: workspace too small");
    return SE3_EWORKSPACE;
  }
  return bq_prepare(pts_src, batch_src, n_src, min_pt, num_cells, radius, w, key_bits, as_stream(stream), n_batches, max_seg);
}

int se3::knn_sorted_fused(const float* pts, const int32_t* batch_ids, int64_t n, const float* raw_min, const float* raw_max,
                          int* idx_sorted, void* pts_sorted_f4, float* minmax, int32_t n_batches, int32_t max_seg,
                          se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 1 && n < (1ll << 31) && seg_build_possible(n_batches, max_seg), "bad arguments");
  SegBuildArgs a;
  a.pts = pts; a.batch = batch_ids; a.n = (int)n; a.n_batches = n_batches;
  a.raw_min = raw_min; a.raw_max = raw_max; a.cell = 1.0f; a.max_pad = 0.0f;
  a.keys_sorted = nullptr; a.idx_sorted = idx_sorted; a.pts_sorted = reinterpret_cast<float4*>(pts_sorted_f4);
  a.mn_out = nullptr; a.mx_out = nullptr; a.nc_out = nullptr; a.mm_out = minmax;
  a.zero = nullptr; a.n_zero = 0;
  return launch_seg_build(1, a, max_seg, as_stream(stream));
}

// sorted source structure of a ball query straight from the raw boxes (one launch); also writes the padded box
// and the grid extents the count pass reads
int se3::ball_query_prepare_fused(const float* pts_src, const int32_t* batch_src, int64_t n_src, const float* raw_min,
                                  const float* raw_max, float radius, float* min_pt_out, float* max_pt_out,
                                  int32_t* num_cells_out, void* ws_src, size_t ws_src_bytes, int32_t n_batches,
                                  int32_t max_seg, se3_stream_t stream) {
  SE3_CHECK_ARG(n_src >= 1 && n_src < (1ll << 31) && radius > 0.0f && seg_build_possible(n_batches, max_seg), "bad arguments");
  BQWorkspace w;
  char dummy[1];
  if (!bq_layout2(ws_src, ws_src_bytes, dummy, (size_t)1 << 60, n_src, 0, 1, w)) {
    set_error("ball_query_prepare_fused: workspace too small");
    return SE3_EWORKSPACE;
  }
  SegBuildArgs a;
  a.pts = pts_src; a.batch = batch_src; a.n = (int)n_src; a.n_batches = n_batches;
  a.raw_min = raw_min; a.raw_max = raw_max; a.cell = radius; a.max_pad = -1e-6f;
  a.keys_sorted = reinterpret_cast<uint64_t*>(w.keys_sorted); a.idx_sorted = w.idx_sorted; a.pts_sorted = w.pts_sorted;
  a.mn_out = min_pt_out; a.mx_out = max_pt_out; a.nc_out = num_cells_out; a.mm_out = nullptr;
  a.zero = nullptr; a.n_zero = 0;
  return launch_seg_build(0, a, max_seg, as_stream(stream));
}

extern "C" int se3_ball_query_count_prepared(const float* pts_dst, const int32_t* batch_dst, int64_t n_src,
                                             int64_t n_dst, int64_t n_dst_max, const float* min_pt,
                                             const int32_t* num_cells, const float* radius, void* ws_src,
                                             size_t ws_src_bytes, void* ws_dst, size_t ws_dst_bytes,
                                             int32_t* row_ends_out, int64_t* total_out, se3_stream_t stream) {
  SE3_CHECK_ARG(n_src >= 0 && n_dst >= 0 && n_src < (1ll << 31) && n_dst < (1ll << 31) && total_out, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (n_dst == 0 || n_src == 0) {
    if (n_dst > 0) SE3_CUDA(cudaMemsetAsync(row_ends_out, 0, n_dst * sizeof(int32_t), st));
    SE3_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    return SE3_OK;
  }
  SE3_CHECK_ARG(pts_dst && batch_dst && min_pt && num_cells && radius && ws_src && ws_dst && row_ends_out, "null pointer");
  BQWorkspace w;
  if (!bq_layout2(ws_src, ws_src_bytes, ws_dst, ws_dst_bytes, n_src, n_dst, n_dst_max < n_dst ? n_dst : n_dst_max, w)) {
    set_error("se3_ball_query_count_prepared: workspace too small");
    return SE3_EWORKSPACE;
  }
  return bq_count(pts_dst, batch_dst, n_src, n_dst, min_pt, num_cells, radius, w, row_ends_out, total_out, st);
}

extern "C" int se3_ball_query_fill_csr_prepared(const float* pts_dst, int64_t n_src, int64_t n_dst, int64_t n_dst_max,
                                                const float* radius, void* ws_src, size_t ws_src_bytes, void* ws_dst,
                                                size_t ws_dst_bytes, const int32_t* row_ends, int64_t n_edges,
                                                int32_t* col_src_out, int32_t* edge_dst_out, se3_stream_t stream) {
  if (n_edges == 0 || n_dst == 0 || n_src == 0) return SE3_OK;
  SE3_CHECK_ARG(pts_dst && radius && ws_src && ws_dst && row_ends && col_src_out && edge_dst_out, "null pointer");
  BQWorkspace w;
  if (!bq_layout2(ws_src, ws_src_bytes, ws_dst, ws_dst_bytes, n_src, n_dst, n_dst_max < n_dst ? n_dst : n_dst_max, w)) {
    set_error("se3_ball_query_fill_csr_prepared: workspace too small");
    return SE3_EWORKSPACE;
  }
  const int blocks = grid_for(n_dst * 32, 256);
  {
    const BqScanArgs qa{pts_dst, nullptr, n_dst, (int)n_src, nullptr, nullptr, radius, w.keys_sorted, w.pts_sorted, w.idx_sorted, w.ranges, w.counts, row_ends, nullptr, col_src_out, edge_dst_out, nullptr, nullptr};
    k_bq_scan<true><<<blocks, 256, 0, as_stream(stream)>>>(qa);
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_ball_query_fill(const float* pts_dst, int64_t n_src, int64_t n_dst, const float* radius,
                                   const void* workspace, size_t workspace_bytes, const int32_t* row_ends,
                                   int64_t n_edges, int64_t* neighbors_out, se3_stream_t stream) {
  if (n_edges == 0 || n_dst == 0 || n_src == 0) return SE3_OK;
  SE3_CHECK_ARG(pts_dst && radius && workspace && row_ends && neighbors_out, "null pointer");
  BQWorkspace w;
  if (!bq_layout(const_cast<void*>(workspace), workspace_bytes, n_src, n_dst, w)) {
    set_error("se3_ball_query_fill: workspace too small");
    return SE3_EWORKSPACE;
  }
  const int blocks = grid_for(n_dst * 32, 256);
  {
    const BqScanArgs qa{pts_dst, nullptr, n_dst, (int)n_src, nullptr, nullptr, radius, w.keys_sorted, w.pts_sorted, w.idx_sorted, w.ranges, w.counts, row_ends, neighbors_out, nullptr, nullptr, nullptr, nullptr};
    k_bq_scan<true><<<blocks, 256, 0, as_stream(stream)>>>(qa);
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

extern "C" int se3_ball_query_fill_csr(const float* pts_dst, int64_t n_src, int64_t n_dst, const float* radius,
                                       const void* workspace, size_t workspace_bytes, const int32_t* row_ends,
                                       int64_t n_edges, int32_t* col_src_out, int32_t* edge_dst_out,
                                       se3_stream_t stream) {
  if (n_edges == 0 || n_dst == 0 || n_src == 0) return SE3_OK;
  SE3_CHECK_ARG(pts_dst && radius && workspace && row_ends && col_src_out && edge_dst_out, "null pointer");
  BQWorkspace w;
  if (!bq_layout(const_cast<void*>(workspace), workspace_bytes, n_src, n_dst, w)) {
    set_error("se3_ball_query_fill_csr: workspace too small");
    return SE3_EWORKSPACE;
  }
  const int blocks = grid_for(n_dst * 32, 256);
  {
    const BqScanArgs qa{pts_dst, nullptr, n_dst, (int)n_src, nullptr, nullptr, radius, w.keys_sorted, w.pts_sorted, w.idx_sorted, w.ranges, w.counts, row_ends, nullptr, col_src_out, edge_dst_out, nullptr, nullptr};
    k_bq_scan<true><<<blocks, 256, 0, as_stream(stream)>>>(qa);
  }
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// ---------------------------------------------------------------------------------------------
// ball query + transposed CSR in one go (fused hierarchy builder)
// ---------------------------------------------------------------------------------------------
// Single-CTA scans (two independent jobs per launch): the arrays here are one cloud level long (<= a few 10^5
// entries), where the device-wide scan's second launch and its tile-state initialisation cost more than the scan.
struct ScanJob {
  const int* in;
  int* out;
  int n;
  int exclusive;
  int64_t* total;
};
__device__ __forceinline__ void scan_small_body(const ScanJob& j) {
  using BS = cub::BlockScan<int, 1024>;
  __shared__ typename BS::TempStorage tmp;
  constexpr int IT = 16;  // 16 K entries per tile: four 128-bit loads in flight per thread (arrays are 16-byte aligned)
  int carry = 0;
  for (int base = 0; base < j.n; base += 1024 * IT) {
    int v[IT];
    const int p0 = base + threadIdx.x * IT;
    if (p0 + IT <= j.n) {
      const int4* src = reinterpret_cast<const int4*>(j.in + p0);
#pragma unroll
      for (int q = 0; q < IT / 4; ++q) {
        const int4 t = src[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < IT; ++i) v[i] = (p0 + i < j.n) ? j.in[p0 + i] : 0;
    }
    int sum = 0;
#pragma unroll
    for (int i = 0; i < IT; ++i) sum += v[i];
    int excl, agg;
    BS(tmp).ExclusiveSum(sum, excl, agg);
    int run = carry + excl;
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const int before = run;
      run += v[i];
      v[i] = j.exclusive ? before : run;
    }
    if (p0 + IT <= j.n) {
      int4* dst = reinterpret_cast<int4*>(j.out + p0);
#pragma unroll
      for (int q = 0; q < IT / 4; ++q) dst[q] = make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < IT; ++i)
        if (p0 + i < j.n) j.out[p0 + i] = v[i];
    }
    carry += agg;
    __syncthreads();
  }
  if (threadIdx.x == 0 && j.total) *j.total = carry;
}
__global__ void __launch_bounds__(1024) k_scan_small(const ScanJob j0, const ScanJob j1) {
  scan_small_body(blockIdx.x == 0 ? j0 : j1);
}

struct ScanBatch {
  ScanJob j[2 * BQ_BATCH];
};
__global__ void __launch_bounds__(1024) k_scan_multi(const ScanBatch b) { scan_small_body(b.j[blockIdx.x]); }

// Every transposed row in ascending edge order (= the order a stable sort by source would give), then the sample
// of every transposed entry.  One warp per row, any row length: chunks of up to T_CHUNK entries are sorted in shared
// memory (bitonic network), longer rows are finished by rank-merging the sorted runs, ping-ponging between the row's
// t_edge slice and its (not yet written) t_dst slice.  Edge ids are distinct, so ranks are unambiguous.
constexpr int T_CHUNK = 2048;
constexpr int T_WARPS = 4;
__device__ __forceinline__ int lower_bound_run(const int* __restrict__ a, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
struct FinishJob {
  const int* t_row_ends;
  int64_t n_src;
  int* t_edge;
  const int* edge_dst;
  int* t_dst;
};
struct FinishBatch {
  FinishJob j[BQ_BATCH];
  int blk[BQ_BATCH + 1];
};
__device__ __forceinline__ void t_rows_finish_body(const int* __restrict__ t_row_ends, int64_t n_src,
                                                   int* __restrict__ t_edge, const int* __restrict__ edge_dst,
                                                   int* __restrict__ t_dst, int block, int nblocks) {
  __shared__ int s_buf[T_WARPS][T_CHUNK];
  const int lane = threadIdx.x & 31;
  int* sb = s_buf[threadIdx.x >> 5];
  const int64_t warp = (block * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)nblocks * blockDim.x) >> 5;
  for (int64_t j = warp; j < n_src; j += nwarps) {
    const int lo = j > 0 ? t_row_ends[j - 1] : 0, hi = t_row_ends[j];
    const int len = hi - lo;
    if (len <= 0) continue;
    int* row = t_edge + lo;
    int* alt = t_dst + lo;
    if (len <= 32) {
      // the common case (a row is a neighbourhood): one entry per lane, rank = number of smaller entries (distinct
      // edge ids), everything in registers
      const int v = lane < len ? row[lane] : INT_MAX;
      int rank = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) rank += __shfl_sync(0xffffffffu, v, j) < v ? 1 : 0;
      __syncwarp();
      if (lane < len) {
        row[rank] = v;
        alt[rank] = edge_dst[v];
      }
      continue;
    }
    for (int c0 = 0; c0 < len; c0 += T_CHUNK) {
      const int cn = min(T_CHUNK, len - c0);
      int P = 32;
      while (P < cn) P <<= 1;
      for (int t = lane; t < P; t += 32) sb[t] = t < cn ? row[c0 + t] : INT_MAX;
      __syncwarp();
      for (int k = 2; k <= P; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
          for (int t = lane; t < (P >> 1); t += 32) {
            const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
            const int a = sb[i], b = sb[i | jj];
            if ((a > b) == ((i & k) == 0)) { sb[i] = b; sb[i | jj] = a; }
          }
          __syncwarp();
        }
      for (int t = lane; t < cn; t += 32) row[c0 + t] = sb[t];
      __syncwarp();
    }
    // rank-merge the sorted runs (only rows longer than one chunk)
    int* src = row;
    int* dst = alt;
    for (int width = T_CHUNK; width < len; width <<= 1) {
      for (int r0 = 0; r0 < len; r0 += 2 * width) {
        const int na = min(width, len - r0), nb = max(0, min(width, len - r0 - width));
        const int* A = src + r0;
        const int* B = src + r0 + na;
        for (int t = lane; t < na; t += 32) dst[r0 + t + lower_bound_run(B, nb, A[t])] = A[t];
        for (int t = lane; t < nb; t += 32) dst[r0 + t + lower_bound_run(A, na, B[t])] = B[t];
      }
      __syncwarp();
      int* tmp = src; src = dst; dst = tmp;
    }
    if (src != row) {
      for (int t = lane; t < len; t += 32) row[t] = src[t];
      __syncwarp();
    }
    for (int t = lane; t < len; t += 32) alt[t] = edge_dst[row[t]];
  }
}
__global__ void __launch_bounds__(T_WARPS * 32) k_t_rows_finish(const int* __restrict__ t_row_ends, int64_t n_src,
                                                                int* __restrict__ t_edge, const int* __restrict__ edge_dst,
                                                                int* __restrict__ t_dst) {
  t_rows_finish_body(t_row_ends, n_src, t_edge, edge_dst, t_dst, (int)blockIdx.x, (int)gridDim.x);
}
__global__ void __launch_bounds__(T_WARPS * 32) k_t_rows_finish_multi(const FinishBatch b) {
  int q = 0;
  while (q + 1 < BQ_BATCH && (int)blockIdx.x >= b.blk[q + 1]) ++q;
  const FinishJob& j = b.j[q];
  t_rows_finish_body(j.t_row_ends, j.n_src, j.t_edge, j.edge_dst, j.t_dst, (int)blockIdx.x - b.blk[q], b.blk[q + 1] - b.blk[q]);
}

namespace se3 {
// count pass of a prepared query; also the histogram of the hit sources.  Outputs: row_ends [n_dst] inclusive,
// t_row [n_src] EXCLUSIVE starts of the transposed rows (bq_fill_transposed turns them into inclusive ends),
// *total_out = E.
int bq_count_transposed(const float* pts_dst, const int32_t* batch_dst, int64_t n_src, int64_t n_dst, int64_t n_dst_max,
                        const float* min_pt, const int32_t* num_cells, const float* radius, void* ws_src,
                        size_t ws_src_bytes, void* ws_dst, size_t ws_dst_bytes, int32_t* row_ends_out, int32_t* t_row_out,
                        int64_t* total_out, se3_stream_t stream) {
  SE3_CHECK_ARG(n_src >= 0 && n_dst >= 0 && n_src < (1ll << 31) && n_dst < (1ll << 31) && total_out, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (n_src > 0) SE3_CUDA(cudaMemsetAsync(t_row_out, 0, n_src * sizeof(int32_t), st));
  if (n_dst == 0 || n_src == 0) {
    if (n_dst > 0) SE3_CUDA(cudaMemsetAsync(row_ends_out, 0, n_dst * sizeof(int32_t), st));
    SE3_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    return SE3_OK;
  }
  SE3_CHECK_ARG(pts_dst && batch_dst && min_pt && num_cells && radius && ws_src && ws_dst && row_ends_out && t_row_out,
                "null pointer");
  BQWorkspace w;
  if (!bq_layout2(ws_src, ws_src_bytes, ws_dst, ws_dst_bytes, n_src, n_dst, n_dst_max < n_dst ? n_dst : n_dst_max, w)) {
    set_error("bq_count_transposed: workspace too small");
    return SE3_EWORKSPACE;
  }
  const int blocks = grid_for(n_dst * 32, 256);
  {
    const BqScanArgs qa{pts_dst, batch_dst, n_dst, (int)n_src, min_pt, num_cells, radius, w.keys_sorted, w.pts_sorted, w.idx_sorted, w.ranges, w.counts, nullptr, nullptr, nullptr, nullptr, t_row_out, nullptr};
    k_bq_scan<false><<<blocks, 256, 0, st>>>(qa);
  }
  SE3_LAUNCH_CHECK();
  ScanJob j0{w.counts, row_ends_out, (int)n_dst, 0, total_out};
  ScanJob j1{t_row_out, t_row_out, (int)n_src, 1, nullptr};
  k_scan_small<<<2, 1024, 0, st>>>(j0, j1);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// fill pass: col_src / edge_dst [E] (the CSR columns and the sample of every edge) and the transposed CSR
// (t_row: exclusive starts in, inclusive ends out; t_edge / t_dst [E], rows in ascending edge order)
int bq_fill_transposed(const float* pts_dst, int64_t n_src, int64_t n_dst, int64_t n_dst_max, const float* radius,
                       void* ws_src, size_t ws_src_bytes, void* ws_dst, size_t ws_dst_bytes, const int32_t* row_ends,
                       int64_t n_edges, int32_t* col_src_out, int32_t* edge_dst_out, int32_t* t_row, int32_t* t_edge,
                       int32_t* t_dst, se3_stream_t stream) {
  if (n_edges == 0 || n_dst == 0 || n_src == 0) return SE3_OK;  // t_row is all zero already
  SE3_CHECK_ARG(pts_dst && radius && ws_src && ws_dst && row_ends && col_src_out && edge_dst_out && t_row && t_edge && t_dst,
                "null pointer");
  BQWorkspace w;
  if (!bq_layout2(ws_src, ws_src_bytes, ws_dst, ws_dst_bytes, n_src, n_dst, n_dst_max < n_dst ? n_dst : n_dst_max, w)) {
    set_error("bq_fill_transposed: workspace too small");
    return SE3_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const int blocks = grid_for(n_dst * 32, 256);
  {
    const BqScanArgs qa{pts_dst, nullptr, n_dst, (int)n_src, nullptr, nullptr, radius, w.keys_sorted, w.pts_sorted, w.idx_sorted, w.ranges, w.counts, row_ends, nullptr, col_src_out, edge_dst_out, t_row, t_edge};
    k_bq_scan<true><<<blocks, 256, 0, st>>>(qa);
  }
  SE3_LAUNCH_CHECK();
  k_t_rows_finish<<<grid_for(n_src * 32, T_WARPS * 32), T_WARPS * 32, 0, st>>>(t_row, n_src, t_edge, edge_dst_out, t_dst);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}
}  // namespace se3

namespace se3 {
// Batched forms of bq_count_transposed / bq_fill_transposed: up to BQ_BATCH prepared queries per launch.  The t_row
// arrays of a batch must lie in one contiguous block [t_row_block, t_row_block + t_row_block_bytes) (one memset).
int bq_count_transposed_batch(const BqBatchItem* it, int n, void* t_row_block, size_t t_row_block_bytes,
                              se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 1 && n <= BQ_BATCH && it, "bad batch");
  cudaStream_t st = as_stream(stream);
  if (t_row_block_bytes) SE3_CUDA(cudaMemsetAsync(t_row_block, 0, t_row_block_bytes, st));
  BqScanBatch qb;
  ScanBatch sb;
  int nq = 0, blocks = 0;
  for (int i = 0; i < n; ++i) {
    const BqBatchItem& q = it[i];
    SE3_CHECK_ARG(q.n_src >= 0 && q.n_dst >= 0 && q.n_src < (1ll << 31) && q.n_dst < (1ll << 31) && q.total_out, "bad sizes");
    if (q.n_dst == 0 || q.n_src == 0) {
      if (q.n_dst > 0) SE3_CUDA(cudaMemsetAsync(q.row_ends, 0, q.n_dst * sizeof(int32_t), st));
      SE3_CUDA(cudaMemsetAsync(q.total_out, 0, sizeof(int64_t), st));
      continue;
    }
    BQWorkspace w;
    if (!bq_layout2(q.ws_src, q.ws_src_bytes, q.ws_dst, q.ws_dst_bytes, q.n_src, q.n_dst, 1, w)) {
      set_error("bq_count_transposed_batch: workspace too small");
      return SE3_EWORKSPACE;
    }
    qb.q[nq] = BqScanArgs{q.pts_dst, q.batch_dst, q.n_dst, (int)q.n_src, q.min_pt, q.num_cells, q.radius, w.keys_sorted,
                          w.pts_sorted, w.idx_sorted, w.ranges, w.counts, nullptr, nullptr, nullptr, nullptr, q.t_row, nullptr};
    sb.j[2 * nq] = ScanJob{w.counts, q.row_ends, (int)q.n_dst, 0, q.total_out};
    sb.j[2 * nq + 1] = ScanJob{q.t_row, q.t_row, (int)q.n_src, 1, nullptr};
    qb.blk[nq] = blocks;
    blocks += grid_for(q.n_dst * 32, 256);
    ++nq;
  }
  if (nq == 0) return SE3_OK;
  for (int i = nq; i <= BQ_BATCH; ++i) qb.blk[i] = blocks;
  k_bq_scan_multi<false><<<blocks, 256, 0, st>>>(qb);
  SE3_LAUNCH_CHECK();
  k_scan_multi<<<2 * nq, 1024, 0, st>>>(sb);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

int bq_fill_transposed_batch(const BqBatchItem* it, int n, se3_stream_t stream) {
  SE3_CHECK_ARG(n >= 1 && n <= BQ_BATCH && it, "bad batch");
  cudaStream_t st = as_stream(stream);
  BqScanBatch qb;
  FinishBatch fb;
  int nq = 0, blocks = 0, fblocks = 0;
  for (int i = 0; i < n; ++i) {
    const BqBatchItem& q = it[i];
    if (q.n_edges == 0 || q.n_dst == 0 || q.n_src == 0) continue;  // t_row is all zero already
    BQWorkspace w;
    if (!bq_layout2(q.ws_src, q.ws_src_bytes, q.ws_dst, q.ws_dst_bytes, q.n_src, q.n_dst, 1, w)) {
      set_error("bq_fill_transposed_batch: workspace too small");
      return SE3_EWORKSPACE;
    }
    qb.q[nq] = BqScanArgs{q.pts_dst, nullptr, q.n_dst, (int)q.n_src, nullptr, nullptr, q.radius, w.keys_sorted, w.pts_sorted,
                          w.idx_sorted, w.ranges, w.counts, q.row_ends, nullptr, q.col_src, q.edge_dst, q.t_row, q.t_edge};
    fb.j[nq] = FinishJob{q.t_row, q.n_src, q.t_edge, q.edge_dst, q.t_dst};
    qb.blk[nq] = blocks;
    fb.blk[nq] = fblocks;
    blocks += grid_for(q.n_dst * 32, 256);
    fblocks += grid_for(q.n_src * 32, T_WARPS * 32);
    ++nq;
  }
  if (nq == 0) return SE3_OK;
  for (int i = nq; i <= BQ_BATCH; ++i) {
    qb.blk[i] = blocks;
    fb.blk[i] = fblocks;
  }
  k_bq_scan_multi<true><<<blocks, 256, 0, st>>>(qb);
  SE3_LAUNCH_CHECK();
  k_t_rows_finish_multi<<<fblocks, T_WARPS * 32, 0, st>>>(fb);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}
}  // namespace se3

// ---------------------------------------------------------------------------------------------
// CSR transpose
// ---------------------------------------------------------------------------------------------
__global__ void k_split_neighbors(const int64_t* __restrict__ nb, int64_t e, int* __restrict__ col_src,
                                  int* __restrict__ iota) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
    col_src[i] = (int)nb[2 * i + 1];
    iota[i] = (int)i;
  }
}

__global__ void k_transposed_rows(const int* __restrict__ src_sorted, const int* __restrict__ t_edge,
                                  const int64_t* __restrict__ nb, int64_t e, int64_t n_src,
                                  int* __restrict__ t_row_ends, int* __restrict__ t_dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += stride) {
    t_dst[i] = (int)nb[2 * (int64_t)t_edge[i]];
  }
  // inclusive end of source row j = upper_bound(src_sorted, j)
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_src; j += stride) {
    int lo = 0, hi = (int)e;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (src_sorted[mid] <= (int)j) lo = mid + 1; else hi = mid;
    }
    t_row_ends[j] = lo;
  }
}

static size_t tr_cub_bytes(int64_t e) { return cub_tmp_bytes(2, e); }

extern "C" size_t se3_csr_transpose_workspace_bytes(int64_t n_edges, int64_t n_src) {
  (void)n_src;
  if (n_edges < 1) n_edges = 1;
  return align_up(n_edges * sizeof(int)) * 2 + align_up(tr_cub_bytes(n_edges)) + 256;
}

extern "C" int se3_csr_transpose(const int64_t* neighbors, int64_t n_edges, int64_t n_src, int64_t n_dst,
                                 void* workspace, size_t workspace_bytes, int32_t* col_src, int32_t* t_row_ends,
                                 int32_t* t_edge, int32_t* t_dst, se3_stream_t stream) {
  (void)n_dst;
  SE3_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31) && n_src >= 0, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (n_edges == 0) {
    if (n_src > 0 && t_row_ends) SE3_CUDA(cudaMemsetAsync(t_row_ends, 0, n_src * sizeof(int), st));
    return SE3_OK;
  }
  SE3_CHECK_ARG(neighbors && workspace && col_src && t_row_ends && t_edge && t_dst, "null pointer");
  Arena ar(workspace, workspace_bytes);
  int* iota = ar.take<int>(n_edges);
  int* src_sorted = ar.take<int>(n_edges);
  size_t cb = tr_cub_bytes(n_edges);
  void* tmp = ar.take<char>(cb);
  if (!ar.ok()) {
    set_error("se3_csr_transpose: workspace too small");
    return SE3_EWORKSPACE;
  }
  k_split_neighbors<<<grid_for(n_edges, 256), 256, 0, st>>>(neighbors, n_edges, col_src, iota);
  SE3_LAUNCH_CHECK();
  int bits = 1;
  while ((1ll << bits) < n_src && bits < 31) ++bits;
  SE3_CUDA(cub::DeviceRadixSort::SortPairs(tmp, cb, col_src, src_sorted, iota, t_edge, (int)n_edges, 0, bits, st));
  count_launch(4);
  const int64_t work = n_edges > n_src ? n_edges : n_src;
  k_transposed_rows<<<grid_for(work, 256), 256, 0, st>>>(src_sorted, t_edge, neighbors, n_edges, n_src, t_row_ends,
                                                        t_dst);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

__global__ void k_iota(int* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int)i;
}

__global__ void k_transposed_rows_i32(const int* __restrict__ src_sorted, const int* __restrict__ t_edge,
                                      const int* __restrict__ edge_dst, int64_t e, int64_t n_src,
                                      int* __restrict__ t_row_ends, int* __restrict__ t_dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += stride) t_dst[i] = edge_dst[t_edge[i]];
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_src; j += stride) {
    int lo = 0, hi = (int)e;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (src_sorted[mid] <= (int)j) lo = mid + 1; else hi = mid;
    }
    t_row_ends[j] = lo;
  }
}

// Transposed CSR straight from the int32 CSR columns (col_src [E], edge_dst [E] = sample of every edge).
extern "C" int se3_csr_transpose_i32(const int32_t* col_src, const int32_t* edge_dst, int64_t n_edges, int64_t n_src,
                                     void* workspace, size_t workspace_bytes, int32_t* t_row_ends, int32_t* t_edge,
                                     int32_t* t_dst, se3_stream_t stream) {
  SE3_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31) && n_src >= 0, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (n_edges == 0) {
    if (n_src > 0 && t_row_ends) SE3_CUDA(cudaMemsetAsync(t_row_ends, 0, n_src * sizeof(int), st));
    return SE3_OK;
  }
  SE3_CHECK_ARG(col_src && edge_dst && workspace && t_row_ends && t_edge && t_dst, "null pointer");
  Arena ar(workspace, workspace_bytes);
  int* iota = ar.take<int>(n_edges);
  int* src_sorted = ar.take<int>(n_edges);
  size_t cb = tr_cub_bytes(n_edges);
  void* tmp = ar.take<char>(cb);
  if (!ar.ok()) {
    set_error("se3_csr_transpose_i32: workspace too small");
    return SE3_EWORKSPACE;
  }
  k_iota<<<grid_for(n_edges, 256), 256, 0, st>>>(iota, n_edges);
  SE3_LAUNCH_CHECK();
  int bits = 1;
  while ((1ll << bits) < n_src && bits < 31) ++bits;
  SE3_CUDA(cub::DeviceRadixSort::SortPairs(tmp, cb, col_src, src_sorted, iota, t_edge, (int)n_edges, 0, bits, st));
  count_launch(4);
  const int64_t work = n_edges > n_src ? n_edges : n_src;
  k_transposed_rows_i32<<<grid_for(work, 256), 256, 0, st>>>(src_sorted, t_edge, edge_dst, n_edges, n_src, t_row_ends,
                                                            t_dst);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// first element of every segment of an int32 array (batch id of a voxel: all its points share it)
__global__ void k_segment_first_i32(const int* __restrict__ x, const int64_t* __restrict__ sorted_ids,
                                    const int* __restrict__ seg_ends, int64_t m, int* __restrict__ out) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < m; s += (int64_t)gridDim.x * blockDim.x) {
    const int lo = s > 0 ? seg_ends[s - 1] : 0;
    out[s] = x[sorted_ids[lo]];
  }
}
extern "C" int se3_segment_first_i32(const int32_t* x, const int64_t* sorted_ids, const int32_t* seg_ends, int64_t m,
                                     int32_t* out, se3_stream_t stream) {
  if (m <= 0) return SE3_OK;
  SE3_CHECK_ARG(x && sorted_ids && seg_ends && out, "null pointer");
  k_segment_first_i32<<<grid_for(m, 256), 256, 0, as_stream(stream)>>>(x, sorted_ids, seg_ends, m, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// one uniformly random member of every segment: pts / batch id of point sorted_ids[start + floor(u * count)]
// (GridSubSample with rnd sampling, pc/GridSubSample.py:36-57)
__global__ void k_segment_pick(const float* __restrict__ pts, const int* __restrict__ batch,
                               const int64_t* __restrict__ sorted_ids, const int* __restrict__ seg_ends, int64_t m,
                               const float* __restrict__ u, float* __restrict__ pts_out, int* __restrict__ batch_out,
                               int64_t* __restrict__ picked) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < m; s += (int64_t)gridDim.x * blockDim.x) {
    const int lo = s > 0 ? seg_ends[s - 1] : 0;
    const int cnt = seg_ends[s] - lo;
    int off = (int)floorf(u[s] * (float)cnt);
    off = min(max(off, 0), cnt - 1);
    const int64_t src = sorted_ids[lo + off];
    pts_out[3 * s] = pts[3 * src];
    pts_out[3 * s + 1] = pts[3 * src + 1];
    pts_out[3 * s + 2] = pts[3 * src + 2];
    batch_out[s] = batch[src];
    if (picked) picked[s] = src;
  }
}
extern "C" int se3_segment_pick(const float* pts, const int32_t* batch, const int64_t* sorted_ids,
                                const int32_t* seg_ends, int64_t m, const float* u, float* pts_out,
                                int32_t* batch_out, int64_t* picked_out, se3_stream_t stream) {
  if (m <= 0) return SE3_OK;
  SE3_CHECK_ARG(pts && batch && sorted_ids && seg_ends && u && pts_out && batch_out, "null pointer");
  k_segment_pick<<<grid_for(m, 256), 256, 0, as_stream(stream)>>>(pts, batch, sorted_ids, seg_ends, m, u, pts_out,
                                                                  batch_out, picked_out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

// ---------------------------------------------------------------------------------------------
// segment pooling (grid average / max pooling of coordinates, features, batch ids)
// ---------------------------------------------------------------------------------------------
__global__ void k_segment_pool(const float* __restrict__ x, int c, const int64_t* __restrict__ sorted_ids,
                               const int* __restrict__ seg_ends, int64_t m, int mode, float* __restrict__ out) {
  // one thread per (segment, channel); segments are short (a voxel's points)
  const int64_t total = m * c;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = t / c;
    const int ch = (int)(t - s * c);
    const int lo = s > 0 ? seg_ends[s - 1] : 0;
    const int hi = seg_ends[s];
    float acc = mode == 0 ? 0.0f : -INFINITY;
    for (int p = lo; p < hi; ++p) {
      const float v = x[sorted_ids[p] * c + ch];
      acc = mode == 0 ? acc + v : fmaxf(acc, v);
    }
    if (mode == 0) acc = hi > lo ? acc / (float)(hi - lo) : 0.0f;
    out[t] = acc;
  }
}

extern "C" int se3_segment_pool_f32(const float* x, int64_t n, int32_t c, const int64_t* sorted_ids,
                                    const int32_t* seg_ends, int64_t m, int32_t mode, float* out,
                                    se3_stream_t stream) {
  (void)n;
  SE3_CHECK_ARG(c >= 1 && m >= 0 && (mode == 0 || mode == 1), "bad arguments");
  if (m == 0) return SE3_OK;
  SE3_CHECK_ARG(x && sorted_ids && seg_ends && out, "null pointer");
  k_segment_pool<<<grid_for(m * c, 256), 256, 0, as_stream(stream)>>>(x, c, sorted_ids, seg_ends, m, mode, out);
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}
