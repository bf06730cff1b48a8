/*
 * se3conv3d_b200 -- C ABI of the B200-native (sm_100a) SE(3) group-convolution hot path.
 *
 * This header is the drop-in boundary.  It replaces the pybind11 module `point_cloud_lib_ops`
 * of the reference (point_cloud_lib/custom_ops/ops_list.cpp:19-26) and adds the fused entry
 * points the rewritten internals use.  Every function:
 *   - is `extern "C"`, takes plain device pointers + sizes + a CUDA stream (cudaStream_t passed
 *     as void*), never a torch type;
 *   - borrows its inputs, writes only into caller-provided outputs / workspace (the caller --
 *     the Python host side -- owns all device memory);
 *   - is asynchronous on `stream` and never synchronises the host;
 *   - returns 0 on success, a negative SE3_E* code otherwise; se3_last_error() gives the text.
 *     (The reference silently returns zeros on unsupported K / D,
 *      custom_ops/feature_aggregation/feat_basis_utils.cuh:35-41.)
 *
 * Index conventions (identical to the reference):
 *   neighbour list   int64 [E,2], column 0 = sample (output) index, column 1 = source (input)
 *                    index, rows grouped by sample in increasing sample order
 *                    (custom_ops/ball_query/ball_query.cu:94-103)
 *   row_ends         int32 [M], INCLUSIVE running end offsets ("start_ids_" in the reference,
 *                    custom_ops/ball_query/store_neighbors.cu:263-286)
 *   features         float32 [N*F, C], row = point*F + frame
 *                    (layers/PNEConvLayerRotEquiv.py:94-104)
 *   frames           float32 [N, F, 9], row-major 3x3 whose COLUMNS are the axes
 *                    (pc/RotationFunctions.py:307-406)
 */
#ifndef SE3CONV3D_B200_H_
#define SE3CONV3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE3_OK 0
#define SE3_EINVAL (-1)    /* bad argument / unsupported shape */
#define SE3_ECUDA (-2)     /* CUDA runtime error (text in se3_last_error) */
#define SE3_EWORKSPACE (-3)/* workspace too small */

typedef void* se3_stream_t; /* cudaStream_t */

int se3_abi_version(void);
const char* se3_last_error(void);
/* number of kernels launched by this library since load (all entry points); for gpu_launches */
int64_t se3_launch_count(void);

/* Optional device timing of the three dominant kernels (CUDA events on the launching stream around every
 * launch; off by default): ids 0 = forward aggregation, 1 = transposed aggregation (data gradient),
 * 2 = edge-gradient kernel.  se3_profile_read synchronises, returns the accumulated milliseconds and launch
 * counts since the last read, and resets.  bench.py uses it for the per-kernel roofline. */
#define SE3_PROF_KERNELS 3
void se3_profile_enable(int32_t on);
int se3_profile_read(double* ms_out, int64_t* count_out);

/* ------------------------------------------------------------------------------------------
 * Grid keys.  Replaces compute_keys (custom_ops/ball_query/compute_keys.cu:76-125, cell/key
 * math custom_ops/ball_query/grid_utils.cuh:56-93; caller custom_ops/ComputeKeys.py:35-40).
 *   cell = clamp(floor((p - aabb_min[b]) * (1/cell_size)), 0, num_cells-1)
 *   key  = ((b*nx + x)*ny + y)*nz + z
 * aabb_min [B,3], num_cells [3], cell_size [3] are DEVICE arrays, as in the reference. */
int se3_compute_keys(const float* pts, const int32_t* batch_ids, int64_t n,
                     const float* aabb_min, const int32_t* num_cells, const float* cell_size,
                     int64_t* keys_out, se3_stream_t stream);

/* Per-batch bounding box and grid extents without host round trips.  Replaces the scatter_min /
 * scatter_max + elementwise chain of pc/BoundingBox.py:17-18 + pc/Grid.py:26-28 (max_pad = +1e-6) and of
 * custom_ops/BallQuery.py:36-39 (max_pad = -1e-6: the wrapper subtracts 1e-6 from the maximum too):
 *   min_pt[b] = min_b(p) - 1e-6,  max_pt[b] = max_b(p) + max_pad,
 *   num_cells = max_b( int((max_pt - min_pt) * (1/cell)) + 1 )      (fp32; torch's CUDA division by a
 *   host scalar is a multiplication by the reciprocal, which is what the reference executes). */
int se3_grid_setup(const float* pts, const int32_t* batch_ids, int64_t n, int32_t n_batches, float cell,
                   float max_pad, float* min_pt_out, float* max_pt_out, int32_t* num_cells_out,
                   se3_stream_t stream);

/* The two halves of se3_grid_setup for callers that build several grids on one cloud: the raw per-batch box
 * (no padding; empty batches keep (+inf, -inf)), and the padded box + extents for one cell size. */
int se3_bbox(const float* pts, const int32_t* batch_ids, int64_t n, int32_t n_batches, float* min_out,
             float* max_out, se3_stream_t stream);
int se3_grid_extents(const float* raw_min, const float* raw_max, int32_t n_batches, float cell, float max_pad,
                     float* min_pt_out, float* max_pt_out, int32_t* num_cells_out, se3_stream_t stream);

/* Dense cell ranks of a voxel grid (pc/Grid.py:39-58: compute_keys -> unique(return_inverse) -> argsort):
 *   cell_ids [N] int64 = rank of the point's key among the distinct keys (sorted-key order),
 *   sorted_ids [N] int64 = stable argsort(cell_ids), cell_ends [N] int32 = inclusive end of every cell
 *   in that order (first *m_out entries valid), m_out = device int64 number of occupied cells.
 *   key_bits: upper bound on the significant bits of the keys (bounds the radix sort), 0 = unknown (64). */
size_t se3_grid_cells_workspace_bytes(int64_t n);
int se3_grid_cells(const float* pts, const int32_t* batch_ids, int64_t n, const float* min_pt,
                   const int32_t* num_cells, float cell, void* workspace, size_t workspace_bytes,
                   int64_t* cell_ids, int64_t* sorted_ids, int32_t* cell_ends, int64_t* m_out,
                   int32_t key_bits, se3_stream_t stream);

/* Keeps n_keep of the n_cand candidate frames of every point after a uniform random permutation
 * (pc/PointcloudRotEquiv.py:148-168); u [N] uniform in [0,1) supplies the randomness (u = NULL keeps
 * the first n_keep).  cand [N,n_cand,9] -> out [N,n_keep,9]. */
int se3_frames_select(const float* cand, const float* u, int64_t n, int32_t n_cand, int32_t n_keep,
                      float* out, se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Ball query.  Replaces ball_query (custom_ops/ball_query/ball_query.cu:22-104 and its five
 * kernels; caller custom_ops/BallQuery.py:44-52).  Two phases because E is data dependent and
 * the caller owns the output allocation:
 *   _count : keys -> sort -> 9 tube ranges per sample -> strict (|s-p|/r < 1) count
 *            -> inclusive row_ends [M] and the total E (device int64 scalar)
 *   _fill  : re-scans the cached ranges and writes neighbours [E,2] int64 in a DETERMINISTIC
 *            order (tube offset, then sorted-key position), unlike the reference's atomic order.
 * `radius` [3] and `num_cells` [3], `min_pt` [B,3] are DEVICE arrays as in the reference. */
size_t se3_ball_query_workspace_bytes(int64_t n_src, int64_t n_dst);
int se3_ball_query_count(const float* pts_src, const float* pts_dst,
                         const int32_t* batch_src, const int32_t* batch_dst,
                         int64_t n_src, int64_t n_dst,
                         const float* min_pt, const int32_t* num_cells, const float* radius,
                         void* workspace, size_t workspace_bytes,
                         int32_t* row_ends_out, int64_t* total_out, se3_stream_t stream);
int se3_ball_query_fill(const float* pts_dst, int64_t n_src, int64_t n_dst, const float* radius,
                        const void* workspace, size_t workspace_bytes,
                        const int32_t* row_ends, int64_t n_edges,
                        int64_t* neighbors_out, se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * CSR helpers for the conv kernels: int32 source column + transposed CSR (edges grouped by
 * SOURCE point, stable in edge id) used by the atomic-free data gradient.
 *   col_src [E] int32; t_row_ends [n_src] int32 inclusive; t_edge [E] int32 (edge id);
 *   t_dst [E] int32 (sample index of that edge). */
size_t se3_csr_transpose_workspace_bytes(int64_t n_edges, int64_t n_src);
int se3_csr_transpose(const int64_t* neighbors, int64_t n_edges, int64_t n_src, int64_t n_dst,
                      void* workspace, size_t workspace_bytes,
                      int32_t* col_src, int32_t* t_row_ends, int32_t* t_edge, int32_t* t_dst,
                      se3_stream_t stream);

/* int32-CSR variants used by the fused hierarchy builder: the fill writes the source column and the
 * sample of every edge as int32 (the [E,2] int64 pair list of the reference contract is then only
 * materialised on demand), and the transposed CSR is built from those two columns. */
int se3_ball_query_fill_csr(const float* pts_dst, int64_t n_src, int64_t n_dst, const float* radius,
                            const void* workspace, size_t workspace_bytes,
                            const int32_t* row_ends, int64_t n_edges,
                            int32_t* col_src_out, int32_t* edge_dst_out, se3_stream_t stream);
/* Prepared variants: one sorted source structure (se3_ball_query_prepare) per (source cloud, radius) serves
 * every query against it; the workspace is split into a source part and a per-query part. */
size_t se3_ball_query_src_workspace_bytes(int64_t n_src, int64_t n_dst_max);
size_t se3_ball_query_dst_workspace_bytes(int64_t n_dst);
int se3_ball_query_prepare(const float* pts_src, const int32_t* batch_src, int64_t n_src, int64_t n_dst_max,
                           const float* min_pt, const int32_t* num_cells, const float* radius,
                           void* ws_src, size_t ws_src_bytes, int32_t key_bits, se3_stream_t stream);
int se3_ball_query_count_prepared(const float* pts_dst, const int32_t* batch_dst, int64_t n_src, int64_t n_dst,
                                  int64_t n_dst_max, const float* min_pt, const int32_t* num_cells,
                                  const float* radius, void* ws_src, size_t ws_src_bytes, void* ws_dst,
                                  size_t ws_dst_bytes, int32_t* row_ends_out, int64_t* total_out,
                                  se3_stream_t stream);
int se3_ball_query_fill_csr_prepared(const float* pts_dst, int64_t n_src, int64_t n_dst, int64_t n_dst_max,
                                     const float* radius, void* ws_src, size_t ws_src_bytes, void* ws_dst,
                                     size_t ws_dst_bytes, const int32_t* row_ends, int64_t n_edges,
                                     int32_t* col_src_out, int32_t* edge_dst_out, se3_stream_t stream);
int se3_csr_transpose_i32(const int32_t* col_src, const int32_t* edge_dst, int64_t n_edges, int64_t n_src,
                          void* workspace, size_t workspace_bytes,
                          int32_t* t_row_ends, int32_t* t_edge, int32_t* t_dst, se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * k-NN (self included, within batch, k <= 64 as in the reference).  Replaces knn_query
 * (custom_ops/knn_query/knn_query.cu:18-197; caller custom_ops/KNNQuery.py:30-33).
 * out [N,k] int32, ascending distance, -1 padded. */
size_t se3_knn_workspace_bytes(int64_t n);
int se3_knn_query(const float* pts, const int32_t* batch_ids, int64_t n, int32_t k,
                  void* workspace, size_t workspace_bytes, int32_t* out, se3_stream_t stream);
/* k-NN of every sample among the sources of ITS batch item, two different clouds (pc/KnnNeighborhood.py:78-84, the
 * torch_cluster.knn branch used by the global-pooling convolutions).  Sources are grouped by batch item:
 * src_batch_ends [B] int32 = inclusive end of item b.  out [M,k] int32: source ids by ascending distance, -1 padded. */
int se3_knn_cross(const float* pts_src, const int32_t* src_batch_ends, const float* pts_dst, const int32_t* batch_dst,
                  int64_t m, int32_t k, int32_t* out, se3_stream_t stream);

/* PCA reference frames from a k-NN table (pc/RotationFunctions.py:307-406).
 * knn [N,k] int32 (-1 => self loop).  fixed_axis: -1 none (4 frames) | 0,1,2 (2 frames; 0 behaves
 * as "none" in the reference because of `not axis_fixed`, pc/RotationFunctions.py:323).
 * frames_out [N, 4 or 2, 9]. */
int se3_pca_frames(const float* pts, const int32_t* knn, int64_t n, int32_t k, int32_t fixed_axis,
                   float* frames_out, se3_stream_t stream);

/* SO(3) frames from normal-distributed quaternions (pc/RotationFunctions.py:176-216, 53-82):
 * q [n,4] (torch.randn output, so RNG parity is kept) -> frames [n,9]. */
int se3_quat_frames(const float* q, int64_t n, float* frames_out, se3_stream_t stream);

/* Segment mean / max over dense cell ids (GridSubSample.__subsample_tensor__,
 * pc/GridSubSample.py:59-72).  sorted_ids [N] = argsort(cell_ids) (stable), seg_ends [M]
 * inclusive ends of each cell in that order.  mode 0 = mean (float32 [N,C]), 1 = max. */
int se3_segment_pool_f32(const float* x, int64_t n, int32_t c, const int64_t* sorted_ids,
                         const int32_t* seg_ends, int64_t m, int32_t mode, float* out,
                         se3_stream_t stream);

/* out[s] = x[first point of cell s] (int32; the batch id of a voxel -- the reference pools batch ids
 * with scatter_max, pc/GridSubSample.py:72, and all points of a voxel share one id). */
int se3_segment_first_i32(const int32_t* x, const int64_t* sorted_ids, const int32_t* seg_ends, int64_t m,
                          int32_t* out, se3_stream_t stream);
/* One uniformly random point per cell (GridSubSample with p_rnd_sample, pc/GridSubSample.py:36-57):
 * picks sorted_ids[start + floor(u[s] * count)]; u [m] uniform in [0,1).  picked_out (optional) [m] int64. */
int se3_segment_pick(const float* pts, const int32_t* batch, const int64_t* sorted_ids,
                     const int32_t* seg_ends, int64_t m, const float* u, float* pts_out,
                     int32_t* batch_out, int64_t* picked_out, se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused hierarchy construction: everything `create_hierarchy`
 * (tasks/SemSeg/train_dfaust_rot.py:108-158) and the model's neighbourhood requests
 * (models/Encoder.py:134-154, Decoder.py:72-80, FPNDecoder.py:104-113) build per step --
 *   level 0 = grid average of the raw cloud (init_cell), n_pool further grid-average levels,
 *   k-NN + PCA frames + random frame selection + gather records for every cloud,
 *   the optional output cloud (one random raw point per init_cell voxel) with its frames,
 *   every requested ball-query neighbourhood as int32 CSR + transposed CSR
 * -- in ONE native call over a caller-provided device arena: (n_pool + 2) blocking size reads instead
 * of one per object, no per-object host allocation.  Same kernels as the per-object entry points.
 * Offsets in the result are BYTE offsets into the arena; 0-sized objects have offset 0. */
#define SE3_HIER_MAX_CLOUDS 10
#define SE3_HIER_MAX_NEIGH 32
typedef struct se3_hier_desc {
  int64_t n;             /* raw input points */
  int32_t n_batches;
  int32_t n_pool;        /* grid-average levels below level 0 */
  float init_cell;       /* level 0 = grid average of the raw cloud with this voxel size (> 0) */
  float cells[SE3_HIER_MAX_CLOUDS]; /* voxel size of pooling step l -> l+1 */
  int32_t knn_k;         /* PCA frames: k nearest neighbours (<= 32); 0 = sampled frames: u_frames holds Gaussian
                            quaternion components [(n_pool + 2) * n * n_frames, 4] instead of uniform variates */
  int32_t n_frames;      /* frames kept per point (1..4) */
  int32_t fixed_axis;    /* -1 none */
  int32_t out_cloud;     /* 1: also build the output cloud (index n_pool + 1) */
  int32_t n_neigh;
  int32_t neigh_src[SE3_HIER_MAX_NEIGH]; /* cloud indices */
  int32_t neigh_dst[SE3_HIER_MAX_NEIGH];
  float neigh_radius[SE3_HIER_MAX_NEIGH];
} se3_hier_desc;
typedef struct se3_hier_cloud {
  int64_t n;
  int64_t pts, batch, frames, rec;         /* f32 [n,3], i32 [n], f32 [n,F,9], f32 [n*F,12] */
  int64_t m;                                /* occupied cells of the grid built on this cloud (0: none) */
  int64_t cell_ids, sorted_ids, cell_ends;  /* i64 [n], i64 [n], i32 [m] */
} se3_hier_cloud;
typedef struct se3_hier_neigh {
  int64_t e;
  int64_t row_ends, col_src, edge_dst, t_row_ends, t_edge, t_dst; /* i32 [n_dst], [e], [e], [n_src], [e], [e] */
} se3_hier_neigh;
typedef struct se3_hier_result {
  int64_t arena_used;       /* bytes consumed (on SE3_EWORKSPACE: bytes wanted by the failing allocation) */
  int32_t n_clouds;
  int32_t reserved;
  se3_hier_cloud raw;       /* the raw cloud's init_cell grid (pts/batch/frames/rec unused) */
  int64_t out_picked;       /* i64 [clouds[out].n]: raw index of every output-cloud point (0 if no out cloud) */
  se3_hier_cloud clouds[SE3_HIER_MAX_CLOUDS];
  se3_hier_neigh neigh[SE3_HIER_MAX_NEIGH];
} se3_hier_result;
/* u_frames: >= (n_pool + 2) * n uniforms (device); u_cells: >= n uniforms (device, only with out_cloud). */
int se3_hierarchy_build(const se3_hier_desc* d, const float* pts, const int32_t* batch_ids,
                        const float* u_frames, const float* u_cells, void* arena, size_t arena_bytes,
                        se3_hier_result* out, se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Legacy aggregation ops (custom_ops/feature_aggregation/feat_basis_proj.cuh:27-31 and
 * feat_basis_proj_grads.cuh:29-34; callers custom_ops/FeatBasisProj.py:36-40, 59-66).
 *   T[m,c,k] = sum_{e in row m} feats[nbr[e,1], c] * basis[e,k]
 * neighbors int32 [E,2]; ends int32 [M] inclusive; K in {8,16,32,64}; any C >= 1.
 * The gradient is atomic-free when `t_*` (transposed CSR over column 1) is given, else uses
 * fp32 atomics like the reference. */
int se3_feat_basis_proj(const float* basis, const float* feats, const int32_t* neighbors,
                        const int32_t* ends, int64_t n_edges, int64_t m, int32_t c, int32_t k,
                        float* out, se3_stream_t stream);
int se3_feat_basis_proj_grad(const float* basis, const float* feats, const int32_t* neighbors,
                             const int32_t* ends, const float* grads, int64_t n_edges, int64_t m,
                             int64_t n_feat_rows, int32_t c, int32_t k, float* feat_grads,
                             float* basis_grads, se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused PNEConvLayerRotEquiv (layers/PNEConvLayerRotEquiv.py:61-128, 160-216).
 * The relative geometry g[e,a,b] (9 values), the basis h = gelu(g.W9 + b) and the expanded
 * neighbour list never reach HBM.
 *
 * precision: 0 = fp32 CUDA-core path (exactness mode, <=1e-4 rel vs the reference)
 *            1 = bf16 tensor-core path (mma.sync aggregation + tcgen05/TMEM projection),
 *                fp32 accumulation everywhere.
 */
typedef struct se3_conv_desc {
  int64_t n_in;        /* input points N */
  int64_t n_out;       /* output points M */
  int64_t n_edges;     /* E */
  int32_t f_in;        /* frames per input point  (1..4) */
  int32_t f_out;       /* frames per output point (1..4) */
  int32_t c_in;        /* input channels */
  int32_t c_out;       /* output channels */
  int32_t k;           /* basis functions (32) */
  int32_t act;         /* 0 linear, 1 relu, 2 gelu(erf), 3 sin */
  int32_t precision;   /* see above */
  int32_t reserved;
  float norm_neigh_dist;   /* buffer norm_neigh_dist_ (layers/IConvLayer.py:33-36) */
  float out_scale;         /* norm_num_neighs_ / f_in (PNEConvLayerRotEquiv.py:213-216) */
  /* geometry (device) */
  const float* pts_in;     /* [N,3] */
  const float* pts_out;    /* [M,3] */
  const float* frames_in;  /* [N,f_in,9] */
  const float* frames_out; /* [M,f_out,9] */
  /* forward CSR (device) */
  const int32_t* row_ends; /* [M] inclusive */
  const int32_t* col_src;  /* [E] */
  /* transposed CSR (device; needed by backward) */
  const int32_t* t_row_ends; /* [N] inclusive */
  const int32_t* t_edge;     /* [E] */
  const int32_t* t_dst;      /* [E] */
  /* packed gather records (device; se3_pack_records): [N*f_in,12] / [M*f_out,12] float32 =
   * (px,py,pz, R[0..8]) per (point, frame), 48 B, 16-byte aligned.  Required by precision 1 (one
   * record = three 128-bit loads per gathered neighbour); ignored by precision 0. */
  const float* rec_in;
  const float* rec_out;
  /* parameters (device) */
  const float* proj_axes;    /* [9,K]   */
  const float* proj_biases;  /* [K]     */
  const float* conv_weights; /* [c_in,K,c_out] */
  /* optional (precision 1): device buffer of se3_conv_weight_cache_bytes() for the bf16 operand layouts of
   * conv_weights.  They depend on the weights only, i.e. they change once per optimiser step, not per call:
   * weight_cache_state 1 = the forward (re)builds them into the buffer, 2 = the buffer is up to date and the forward
   * skips the conversion (evaluation, gradient accumulation, several calls per step).  NULL / 0 = rebuilt on every
   * forward inside the call's saved buffer.  The backward of a call reads the buffer the forward used. */
  void* weight_cache;
  int32_t weight_cache_state;
  int32_t reserved2;
} se3_conv_desc;

/* rec [n*f,12] = (pts[i,0..2], frames[i,a,0..8]) -- the gather record of the tensor-core path. */
int se3_pack_records(const float* pts, const float* frames, int64_t n, int32_t f, float* rec,
                     se3_stream_t stream);

size_t se3_conv_fwd_workspace_bytes(const se3_conv_desc* d);
size_t se3_conv_bwd_workspace_bytes(const se3_conv_desc* d);
/* bytes of the per-call "saved" buffer written by fwd and consumed by bwd (may be 0) */
size_t se3_conv_saved_bytes(const se3_conv_desc* d);
/* bytes of the optional per-layer weight-layout cache (0 for precision 0) */
size_t se3_conv_weight_cache_bytes(const se3_conv_desc* d);

/* y [M*f_out, c_out] = conv(x [N*f_in, c_in]) */
int se3_conv_fwd(const se3_conv_desc* d, const float* x, float* y, void* saved,
                 void* workspace, size_t workspace_bytes, se3_stream_t stream);
/* gradients of y wrt x, conv_weights, proj_axes, proj_biases; any output may be NULL */
int se3_conv_bwd(const se3_conv_desc* d, const float* x, const float* dy, const void* saved,
                 float* dx, float* d_conv_weights, float* d_proj_axes, float* d_proj_biases,
                 void* workspace, size_t workspace_bytes, se3_stream_t stream);

/* ---- the glue around the convolution ---------------------------------------------------
 * out = drop_path(x * gamma) + y: layers/SkipConnection.py:31-43 with layers/DropPathPC.py:23-50.  x, y, out [rows, c];
 * gamma [c]; item_scale [B] = keep_mask / keep_prob per batch item (NULL = no drop path); point_item [rows / frames] =
 * batch item of every POINT (rows are (point, frame), `frames` rows per point).  bwd: dx (may be NULL) and dgamma [c]
 * (ordered two-stage reduction through `workspace`); the gradient of y is dy itself. */
size_t se3_gamma_skip_workspace_bytes(int64_t rows, int32_t c);
int se3_gamma_skip_fwd(const float* x, const float* y, const float* gamma, const float* item_scale,
                       const int32_t* point_item, int32_t frames, int64_t rows, int32_t c, float* out, se3_stream_t stream);
int se3_gamma_skip_bwd(const float* dy, const float* x, const float* gamma, const float* item_scale,
                       const int32_t* point_item, int32_t frames, int64_t rows, int32_t c, float* dx, float* dgamma,
                       void* workspace, size_t workspace_bytes, se3_stream_t stream);
/* pooling of the f per-frame rows of every point (pc/PointcloudRotEquiv.py:224-251): x [n*f, c] -> out [n, c];
 * mode 0 avg, 1 sum, 2 max, 3 min.  bwd needs x and out for max / min (first frame attaining the extremum). */
int se3_frame_pool_fwd(const float* x, int64_t n, int32_t f, int32_t c, int32_t mode, float* out, se3_stream_t stream);
int se3_frame_pool_bwd(const float* dout, const float* x, const float* out, int64_t n, int32_t f, int32_t c, int32_t mode,
                       float* dx, se3_stream_t stream);
/* pooling of the rows of every batch item (global_pooling*, pc/PointcloudRotEquiv.py:195-222, 253-275): rows are grouped
 * by item, item_ends [B] int32 inclusive ends; modes as above.  bwd: mode 0 / 1, or 4 = plain gather of dout rows
 * (global_upsample, :277-286); row_item [rows] = item of every row. */
int se3_batch_pool_fwd(const float* x, const int32_t* item_ends, int32_t n_items, int32_t c, int32_t mode, float* out,
                       se3_stream_t stream);
int se3_batch_pool_bwd(const float* dout, const int32_t* item_ends, const int32_t* row_item, int64_t rows, int32_t c,
                       int32_t mode, float* dx, se3_stream_t stream);

/* Kernel selection of precision 1 for layers with 9..64 gathered channels and 1..2 row frames:
 *   0 (default) aggregation kernel (mma.sync) + projection GEMM (tcgen05), T [R, Cin*K] through HBM;
 *   1 warp-specialised tcgen05 aggregation kernel (conv_fused.cu) + projection GEMM;
 *   2 aggregation AND projection in the fused tcgen05 kernel when the projection weights fit in shared memory (T stays
 *     on chip; a bf16 copy is still written for the weight gradient).
 * Returns the previous mode.  Byte counts of a descriptor (se3_conv_*_bytes) depend on the mode: query them after. */
int se3_conv_set_fused(int32_t mode);

/* The projection GEMM on its own: C[M,N] = alpha * A[M,K] . B[N,K]^T, bf16 operands stored K-major
 * (row-major [rows][K]), fp32 accumulation, C fp32 or bf16 row-major.  This is the [K*Cin] x Cout
 * contraction of the layer (layers/PNEConvLayerRotEquiv.py:210) and its data-gradient twins.
 * impl: 0 = auto (the persistent TMA-fed tcgen05 kernel when N % 16 == 0, else mma.sync), 1 = mma.sync, 2 = tcgen05 with
 * cp.async operand loads (round-1 kernel), 3 = tcgen05 with TMA operand loads. */
int se3_gemm_bf16_tn(const void* a_bf16, const void* b_bf16, int64_t m, int64_t n, int64_t k, float alpha,
                     void* c, int32_t c_is_bf16, int32_t impl, se3_stream_t stream);

/* The weight-gradient GEMM on its own: C[M,N] = alpha * A^T . B with A stored [K][M] and B stored [K][N] (bf16,
 * row-major: the contraction index is the row index of both operands), fp32 accumulation and output.  This is
 * dW = T^T dy of the layer's backward (the autograd twin of layers/PNEConvLayerRotEquiv.py:210).  M and N multiples of 8.
 * impl: 0 = auto (the persistent TMA-fed tcgen05 kernel when the operands are 16-byte aligned), 1 = mma.sync,
 * 2 = tcgen05 with cp.async operand loads (round-1 kernel), 3 = tcgen05 with TMA operand loads. */
int se3_gemm_bf16_mn(const void* a_bf16, const void* b_bf16, int64_t m, int64_t n, int64_t k, float alpha, float* c,
                     int32_t impl, se3_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SE3CONV3D_B200_H_ */
