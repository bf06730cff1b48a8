// Shared helpers for the se3conv3d_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/se3conv3d_b200.h"

namespace se3 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// optional per-kernel timing (se3_profile_enable / se3_profile_read)
bool profile_enabled();
void profile_begin(int id, cudaStream_t st, void** handle);
void profile_end(void* handle, cudaStream_t st);
struct ProfScope {
  void* h;
  cudaStream_t st;
  ProfScope(int id, cudaStream_t s) : st(s) { profile_begin(id, s, &h); }
  ~ProfScope() { profile_end(h, st); }
};

// key sort shared by the grid / ball-query / kNN builders (grid_ops.cu): per-batch-item CTA sort when
// 0 < max_seg <= 7168 points per item, device-wide radix sort otherwise
constexpr int kSegSortMax = 7168;
size_t cub_tmp_bytes(int kind, int64_t n);
size_t sort_pairs_tmp_bytes(int64_t n);
int batch_counts(const int32_t* batch_ids, int64_t n, int32_t n_batches, int32_t* counts_out, cudaStream_t st);
int sort_keys_u64(const uint64_t* keys_in, const int* iota, const int* batch, int64_t n, int n_batches, int max_seg,
                  uint64_t* keys_out, int* idx_out, int end_bit, void* cub_tmp, size_t cub_bytes, cudaStream_t st);
// entry points with the batch-size hint of the fused hierarchy builder (the extern "C" versions pass 0 = unknown)
int grid_cells_impl(const float* pts, const int32_t* batch_ids, int64_t n, const float* min_pt, const int32_t* num_cells,
                    float cell, void* workspace, size_t workspace_bytes, int64_t* cell_ids, int64_t* sorted_ids,
                    int32_t* cell_ends, int64_t* m_out, int32_t key_bits, int32_t n_batches, int32_t max_seg,
                    int32_t* batch_cells, se3_stream_t stream);
int ball_query_prepare_impl(const float* pts_src, const int32_t* batch_src, int64_t n_src, int64_t n_dst_max,
                            const float* min_pt, const int32_t* num_cells, const float* radius, void* ws_src,
                            size_t ws_src_bytes, int32_t key_bits, int32_t n_batches, int32_t max_seg, se3_stream_t stream);
int bq_count_transposed(const float* pts_dst, const int32_t* batch_dst, int64_t n_src, int64_t n_dst, int64_t n_dst_max,
                        const float* min_pt, const int32_t* num_cells, const float* radius, void* ws_src,
                        size_t ws_src_bytes, void* ws_dst, size_t ws_dst_bytes, int32_t* row_ends_out, int32_t* t_row_out,
                        int64_t* total_out, se3_stream_t stream);
int bq_fill_transposed(const float* pts_dst, int64_t n_src, int64_t n_dst, int64_t n_dst_max, const float* radius,
                       void* ws_src, size_t ws_src_bytes, void* ws_dst, size_t ws_dst_bytes, const int32_t* row_ends,
                       int64_t n_edges, int32_t* col_src_out, int32_t* edge_dst_out, int32_t* t_row, int32_t* t_edge,
                       int32_t* t_dst, se3_stream_t stream);
bool seg_build_possible(int n_batches, int max_seg);
int grid_level_fused(const float* pts, const int32_t* batch_ids, int64_t n, const float* raw_min, const float* raw_max,
                     float cell, float* min_pt_out, float* max_pt_out, int32_t* num_cells_out, int64_t* cell_ids,
                     int64_t* sorted_ids, int32_t* cell_ends, int64_t* m_out, int32_t n_batches, int32_t max_seg,
                     int32_t* batch_cells, int32_t* state, float* pool_pts, int32_t* pool_batch, float* pool_min,
                     float* pool_max, se3_stream_t stream);
int ball_query_prepare_fused(const float* pts_src, const int32_t* batch_src, int64_t n_src, const float* raw_min,
                             const float* raw_max, float radius, float* min_pt_out, float* max_pt_out,
                             int32_t* num_cells_out, void* ws_src, size_t ws_src_bytes, int32_t n_batches,
                             int32_t max_seg, se3_stream_t stream);
// kNN sweep structure in one launch (grid_ops.cu): idx_sorted / pts_sorted (xyz + batch bits) / minmax [6]
constexpr int kBqBatch = 16;  // prepared ball queries per batched launch (bounded by the kernel parameter space)
// one prepared ball query of a batched count / fill launch (fused hierarchy builder)
struct BqBatchItem {
  const float* pts_dst;
  const int32_t* batch_dst;
  int64_t n_src, n_dst;
  const float* min_pt;
  const int32_t* num_cells;
  const float* radius;
  void* ws_src;
  size_t ws_src_bytes;
  void* ws_dst;
  size_t ws_dst_bytes;
  int32_t* row_ends;   // [n_dst] inclusive
  int32_t* t_row;      // [n_src] transposed rows: exclusive starts after the count, inclusive ends after the fill
  int64_t* total_out;  // device: E
  int64_t n_edges;     // fill: E (host)
  int32_t *col_src, *edge_dst, *t_edge, *t_dst;  // fill outputs [E]
};
int bq_count_transposed_batch(const BqBatchItem* items, int n, void* t_row_block, size_t t_row_block_bytes,
                              se3_stream_t stream);
int bq_fill_transposed_batch(const BqBatchItem* items, int n, se3_stream_t stream);
int knn_sorted_fused(const float* pts, const int32_t* batch_ids, int64_t n, const float* raw_min, const float* raw_max,
                     int* idx_sorted, void* pts_sorted_f4, float* minmax, int32_t n_batches, int32_t max_seg,
                     se3_stream_t stream);
int pca_frames_select_pack(const float* pts, const int32_t* knn, int64_t n, int32_t k, int32_t fixed_axis, const float* u,
                           int32_t n_keep, float* frames_out, float* rec_out, se3_stream_t stream);
int knn_query_impl(const float* pts, const int32_t* batch_ids, int64_t n, int32_t k, void* workspace,
                   size_t workspace_bytes, int32_t* out, int32_t n_batches, int32_t max_seg, se3_stream_t stream,
                   const float* raw_min = nullptr, const float* raw_max = nullptr);

inline cudaStream_t as_stream(se3_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define SE3_CHECK_ARG(cond, msg)                     \
  do {                                               \
    if (!(cond)) {                                   \
      se3::set_error("%s: %s", __func__, msg);       \
      return SE3_EINVAL;                             \
    }                                                \
  } while (0)

#define SE3_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      se3::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SE3_ECUDA;                                                                 \
    }                                                                                   \
  } while (0)

#define SE3_LAUNCH_CHECK()                    \
  do {                                        \
    se3::count_launch();                      \
    SE3_CUDA(cudaGetLastError());             \
  } while (0)

// Opt a kernel instantiation into its dynamic shared-memory size once per process (the attribute call costs
// several microseconds; `kern` must name ONE instantiation at the call site, the flag is per call site).
#define SE3_SMEM_ONCE(kern, bytes)                                                                       \
  do {                                                                                                   \
    static size_t _se3_smem_set = 0;                                                                     \
    if ((size_t)(bytes) > _se3_smem_set) {                                                               \
      SE3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
      _se3_smem_set = (size_t)(bytes);                                                                   \
    }                                                                                                    \
  } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Programmatic dependent launch (the kernels of one convolution call form chains of short launches): the grid may be
// scheduled while the previous kernel of the stream drains, which hides the launch latency behind that kernel's tail.
// A kernel launched this way MUST call pdl_wait() before it reads or writes anything in global memory (the wait
// returns once the previous grid has completed and its writes are visible); only on-chip set-up may precede it.
// SE3_PDL=0 launches the same kernels without the attribute.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next kernel of the stream start being scheduled (it still waits for this grid to finish in its own
// pdl_wait).  Used by the persistent kernels, whose CTAs are all resident from the start: the dependent grid then
// fills the SMs as these CTAs retire instead of after the last one.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t c) : base(reinterpret_cast<char*>(p)), cap(c), off(0) {}
  template <typename T>
  T* take(size_t n) {
    size_t bytes = align_up(n * sizeof(T));
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// activation of the point-neighbourhood embedding (layers/PNEConvLayer.py:90-101)
__device__ __forceinline__ float pne_act(float x, int act) {
  switch (act) {
    case 1: return fmaxf(x, 0.0f);
    case 2: return gelu_erf(x);
    case 3: return sinf(x);
    default: return x;
  }
}
__device__ __forceinline__ float pne_act_grad(float x, int act) {
  switch (act) {
    case 1: return x > 0.0f ? 1.0f : 0.0f;
    case 2: return gelu_erf_grad(x);
    case 3: return cosf(x);
    default: return 1.0f;
  }
}

}  // namespace se3
