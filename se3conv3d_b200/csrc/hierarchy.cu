// Fused hierarchy construction: one native call builds every cloud, frame set, gather record and
// ball-query CSR a training step needs (see include/se3conv3d_b200.h, se3_hierarchy_build).  It drives the
// same kernels as the per-object entry points; what it removes is the host work between them: one Python
// object, several tensor allocations and one blocking size read per grid / neighbourhood become
// (n_pool + 2) size reads in total and a bump allocator over a caller-provided arena.
#include <string.h>
#include <stdlib.h>
#include <chrono>
#include <mutex>
#include "common.cuh"

namespace se3 {
namespace {

struct Bump {
  char* base;
  size_t cap, off, want;
  bool ok;
  Bump(void* p, size_t c) : base(reinterpret_cast<char*>(p)), cap(c), off(0), want(0), ok(true) {}
  // returns the byte offset of a fresh block (0-sized requests still get a valid aligned offset)
  int64_t take(size_t bytes) {
    const size_t b = align_up(bytes ? bytes : 1);
    const size_t at = off;
    off += b;
    if (off > cap) {
      ok = false;
      want = off;
      return 0;
    }
    return (int64_t)at;
  }
  template <typename T>
  T* at(int64_t o) const { return reinterpret_cast<T*>(base + o); }
};

constexpr size_t kPinnedBytes = 65536;  // edge totals (<= 32 x 8 B) or the per-batch input boxes (B x 24 B)

// Side streams of the builder: the pooling chain (grid -> pooled cloud -> next grid) is the critical path and stays
// on the caller's stream; frames (kNN + PCA), ball-query sources, counts, fills and transposes of different clouds /
// neighbourhoods are independent and fan out over kSide streams, ordered by events and joined back into the
// caller's stream before every host read and before the call returns.
constexpr int kSide = 4;
struct Lanes {
  bool ready = false;
  std::mutex mu;            // one build per device at a time: the streams, events and the host buffer are shared
  int64_t* host = nullptr;  // pinned buffer of the blocking size reads (kPinnedBytes)
  cudaStream_t s[kSide];
  cudaEvent_t fork, join[kSide], cloud[SE3_HIER_MAX_CLOUDS + 1], frames[SE3_HIER_MAX_CLOUDS + 1], src[SE3_HIER_MAX_NEIGH];
};
Lanes* lanes_for_device() {
  static Lanes all[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  Lanes& l = all[dev];
  static std::mutex init_mu;
  std::lock_guard<std::mutex> init_lock(init_mu);
  if (!l.ready) {
    auto ev = [](cudaEvent_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess; };
    bool ok = ev(&l.fork);
    ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&l.host), kPinnedBytes, cudaHostAllocDefault) == cudaSuccess;
    // the builder's kernels are tiny and latency-bound: on the highest stream priority they are dispatched ahead of queued
    // convolution kernels whenever the build overlaps a conv stack (input pipeline on a second stream)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (int i = 0; i < kSide; ++i)
      ok = ok && cudaStreamCreateWithPriority(&l.s[i], cudaStreamNonBlocking, prio_hi) == cudaSuccess && ev(&l.join[i]);
    for (int i = 0; i <= SE3_HIER_MAX_CLOUDS; ++i) ok = ok && ev(&l.cloud[i]) && ev(&l.frames[i]);
    for (int i = 0; i < SE3_HIER_MAX_NEIGH; ++i) ok = ok && ev(&l.src[i]);
    if (!ok) return nullptr;
    l.ready = true;
  }
  return &l;
}

}  // namespace
}  // namespace se3

using namespace se3;

#define HB_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != SE3_OK) return _rc; \
  } while (0)
#define HB_CHECK_ARENA(what)                                        \
  do {                                                              \
    if (!ar.ok) {                                                   \
      out->arena_used = (int64_t)ar.want;                           \
      set_error("se3_hierarchy_build: arena too small (%s)", what); \
      return SE3_EWORKSPACE;                                        \
    }                                                               \
  } while (0)

extern "C" int se3_hierarchy_build(const se3_hier_desc* d, const float* pts, const int32_t* batch_ids,
                                   const float* u_frames, const float* u_cells, void* arena, size_t arena_bytes,
                                   se3_hier_result* out, se3_stream_t stream) {
  SE3_CHECK_ARG(d && out && arena, "null pointer");
  SE3_CHECK_ARG(d->n >= 1 && d->n < (1ll << 31) && d->n_batches >= 1, "bad sizes");
  SE3_CHECK_ARG(d->n_pool >= 0 && d->n_pool + 2 <= SE3_HIER_MAX_CLOUDS, "too many levels");
  SE3_CHECK_ARG(d->init_cell > 0.0f, "init_cell must be positive");
  SE3_CHECK_ARG(d->knn_k >= 0 && d->knn_k <= 32 && d->n_frames >= 1 && d->n_frames <= 4, "bad frame configuration");
  // knn_k == 0: sampled (Monte-Carlo) frames -- u_frames then holds Gaussian quaternion components, 4 per (point, frame)
  const bool mc_frames = d->knn_k == 0;
  SE3_CHECK_ARG(!mc_frames || d->fixed_axis <= 0, "sampled frames about a fixed axis are not fused");
  SE3_CHECK_ARG(d->n_neigh >= 0 && d->n_neigh <= SE3_HIER_MAX_NEIGH, "too many neighbourhoods");
  SE3_CHECK_ARG(pts && batch_ids && u_frames, "null input");
  SE3_CHECK_ARG(!d->out_cloud || u_cells, "the output cloud needs u_cells");
  cudaStream_t st = as_stream(stream);
  // SE3_HIER_TRACE=1: host time stamps of the build phases on stderr (debugging aid)
  static const bool trace = getenv("SE3_HIER_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto stamp = [&](const char* what) {
    if (trace)
      fprintf(stderr, "[se3_hierarchy_build] %-28s %8.1f us\n", what,
              std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count());
  };
  // per-device side streams, events and the pinned buffer of the size reads (see Lanes)
  Lanes* ln = lanes_for_device();
  if (!ln) {
    set_error("se3_hierarchy_build: cannot create the side streams / pinned host buffer");
    return SE3_ECUDA;
  }
  std::lock_guard<std::mutex> build_lock(ln->mu);
  int64_t* host = ln->host;
  memset(out, 0, sizeof(*out));
  Bump ar(arena, arena_bytes);
  const int n_cand = d->fixed_axis > 0 ? 2 : 4;
  const int F = d->n_frames;
  SE3_CHECK_ARG(mc_frames || F <= n_cand, "n_frames exceeds the PCA candidates");
  const int n_clouds = d->n_pool + 1 + (d->out_cloud ? 1 : 0);
  out->n_clouds = n_clouds;

  // small device scalars: padded min/max [B,3] + num_cells [3] of the grid being built, counters, radii
  const size_t bb = (size_t)d->n_batches * 3 * 4;
  const int64_t o_min = ar.take(bb), o_max = ar.take(bb);
  const int64_t o_nc = ar.take(16), o_cnt = ar.take(8 * (SE3_HIER_MAX_NEIGH + 2)), o_rad = ar.take(16 * SE3_HIER_MAX_NEIGH);
  // [int64 m][int32 per-batch-item sizes of the pooled cloud]: one read per grid
  const size_t lvl_bytes = 8 + (size_t)d->n_batches * 4;
  const int64_t o_lvl = ar.take(lvl_bytes);
  int64_t o_rawmin[SE3_HIER_MAX_CLOUDS + 1], o_rawmax[SE3_HIER_MAX_CLOUDS + 1];  // raw boxes: [0] input cloud, [1+c] cloud c
  for (int c = 0; c <= n_clouds; ++c) {
    o_rawmin[c] = ar.take(bb);
    o_rawmax[c] = ar.take(bb);
  }
  HB_CHECK_ARENA("scalars");
  float* min_pt = ar.at<float>(o_min);
  float* max_pt = ar.at<float>(o_max);
  int32_t* num_cells = ar.at<int32_t>(o_nc);
  int64_t* d_cnt = ar.at<int64_t>(o_cnt);
  int64_t* d_m = ar.at<int64_t>(o_lvl);
  int32_t* d_items = reinterpret_cast<int32_t*>(ar.at<char>(o_lvl) + 8);
  const bool track_items = lvl_bytes <= kPinnedBytes && (size_t)d->n_batches * 28 <= kPinnedBytes;

  // ---- bounding box of the input cloud on the host (one early blocking read): every later cloud lies inside
  // it, so the number of significant key bits of every grid is known up front and bounds the radix sorts.
  HB_TRY(se3_bbox(pts, batch_ids, d->n, d->n_batches, ar.at<float>(o_rawmin[0]), ar.at<float>(o_rawmax[0]), stream));
  if (track_items) HB_TRY(batch_counts(batch_ids, d->n, d->n_batches, d_items, st));
  float ext[3] = {0.f, 0.f, 0.f};
  // largest batch item of the cloud about to be sorted (0 = unknown -> device-wide sorts); pooled clouds are
  // re-measured at every grid read, so small levels get the small per-item sort
  auto max_item = [&](const int32_t* h, int nb) -> int {
    int mx = 0;
    for (int b = 0; b < nb; ++b) mx = h[b] > mx ? h[b] : mx;
    return mx;
  };
  int seg_raw = 0;
  {
    float* hf = reinterpret_cast<float*>(host);
    // large batch counts: fall back to full-width sorts instead of a bigger host buffer
    if ((size_t)d->n_batches * 6 * sizeof(float) <= kPinnedBytes) {
      SE3_CUDA(cudaMemcpyAsync(hf, ar.at<float>(o_rawmin[0]), bb, cudaMemcpyDeviceToHost, st));
      SE3_CUDA(cudaMemcpyAsync(hf + d->n_batches * 3, ar.at<float>(o_rawmax[0]), bb, cudaMemcpyDeviceToHost, st));
      if (track_items)
        SE3_CUDA(cudaMemcpyAsync(hf + d->n_batches * 6, d_items, (size_t)d->n_batches * 4, cudaMemcpyDeviceToHost, st));
      SE3_CUDA(cudaStreamSynchronize(st));
      stamp("input box read");
      if (track_items) seg_raw = max_item(reinterpret_cast<const int32_t*>(hf + d->n_batches * 6), d->n_batches);
      for (int b = 0; b < d->n_batches; ++b)
        for (int k = 0; k < 3; ++k) {
          const float lo = hf[3 * b + k], hi = hf[d->n_batches * 3 + 3 * b + k];
          if (lo <= hi && hi - lo > ext[k]) ext[k] = hi - lo;
        }
    } else {
      ext[0] = ext[1] = ext[2] = -1.0f;
    }
  }
  auto key_bits_for = [&](float cell) -> int {
    if (ext[0] < 0.0f) return 0;
    double keys = (double)d->n_batches;
    for (int k = 0; k < 3; ++k) keys *= (double)((int64_t)(ext[k] / cell) + 3);  // +1 exact, +2 padding / rounding slack
    int bits = 1;
    while (bits < 63 && (double)(1ull << bits) < keys) ++bits;
    return bits;
  };

  // grid on (p, b, n) with voxel `cell` from the cloud's raw box; fills g.cell_ids / sorted_ids / cell_ends / m
  // (blocking read of m)
  int seg_next = 0;  // largest batch item of the cloud the last grid pooled
  // grid on (p, b, n) with voxel `cell` from the cloud's raw box; fills g.cell_ids / sorted_ids / cell_ends.
  // grid_launch issues the kernels and the read of (m, per-item sizes); grid_finish blocks on it.  Whatever the
  // host issues in between (the side-stream work of the previous cloud) overlaps the grid chain on the GPU.
  bool level_fused = false;  // the last grid_launch also produced the pooled cloud and its raw boxes
  auto grid_launch = [&](const float* p, const int32_t* b, int64_t n, int raw_slot, float cell, se3_hier_cloud& g,
                         int max_seg, se3_hier_cloud& dst, int dst_raw_slot) -> int {
    g.cell_ids = ar.take((size_t)n * 8);
    g.sorted_ids = ar.take((size_t)n * 8);
    g.cell_ends = ar.take((size_t)n * 4);
    level_fused = n > 0 && seg_build_possible(d->n_batches, max_seg);
    if (level_fused) {
      // every batch item fits a CTA: the whole level (extents, keys, sort, ranks, pooled cloud, its boxes) is one
      // launch; the pooled arrays are taken at their upper bound (the source size), their length is read below
      dst.pts = ar.take((size_t)n * 12);
      dst.batch = ar.take((size_t)n * 4);
      const int64_t o_state = ar.take((size_t)d->n_batches * 4);
      HB_CHECK_ARENA("grid");
      HB_TRY(grid_level_fused(p, b, n, ar.at<float>(o_rawmin[raw_slot]), ar.at<float>(o_rawmax[raw_slot]), cell, min_pt,
                              max_pt, num_cells, ar.at<int64_t>(g.cell_ids), ar.at<int64_t>(g.sorted_ids),
                              ar.at<int32_t>(g.cell_ends), d_m, d->n_batches, max_seg, track_items ? d_items : nullptr,
                              ar.at<int32_t>(o_state), ar.at<float>(dst.pts), ar.at<int32_t>(dst.batch),
                              ar.at<float>(o_rawmin[dst_raw_slot]), ar.at<float>(o_rawmax[dst_raw_slot]), stream));
    } else {
      const size_t wsb = se3_grid_cells_workspace_bytes(n);
      const int64_t o_ws = ar.take(wsb);
      HB_CHECK_ARENA("grid");
      HB_TRY(se3_grid_extents(ar.at<float>(o_rawmin[raw_slot]), ar.at<float>(o_rawmax[raw_slot]), d->n_batches, cell,
                              1e-6f, min_pt, max_pt, num_cells, stream));
      HB_TRY(grid_cells_impl(p, b, n, min_pt, num_cells, cell, ar.at<char>(o_ws), wsb, ar.at<int64_t>(g.cell_ids),
                             ar.at<int64_t>(g.sorted_ids), ar.at<int32_t>(g.cell_ends), d_m, key_bits_for(cell),
                             d->n_batches, max_seg, track_items ? d_items : nullptr, stream));
    }
    SE3_CUDA(cudaMemcpyAsync(host, d_m, track_items ? lvl_bytes : sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    return SE3_OK;
  };
  auto grid_finish = [&](int64_t n, se3_hier_cloud& g) -> int {
    SE3_CUDA(cudaStreamSynchronize(st));
    g.m = host[0];
    stamp("grid size read");
    seg_next = track_items && n > 0 ? max_item(reinterpret_cast<const int32_t*>(host + 1), d->n_batches) : 0;
    return SE3_OK;
  };
  // raw box of cloud c (on the caller's stream: the next grid of the pooling chain needs it)
  auto cloud_bbox = [&](se3_hier_cloud& c, int raw_slot, se3_stream_t s) -> int {
    return se3_bbox(ar.at<float>(c.pts), ar.at<int32_t>(c.batch), c.n, d->n_batches, ar.at<float>(o_rawmin[raw_slot]),
                    ar.at<float>(o_rawmax[raw_slot]), s);
  };
  // frames + records of cloud c (pts / batch already in place) on stream s
  int64_t u_off = 0;
  int cloud_seg[SE3_HIER_MAX_CLOUDS + 1] = {0};  // largest batch item per cloud (0 = unknown)
  auto build_frames = [&](se3_hier_cloud& c, int max_seg, int raw_slot, se3_stream_t s) -> int {
    const int64_t n = c.n;
    c.frames = ar.take((size_t)n * F * 36);
    c.rec = ar.take((size_t)n * F * 48);
    if (mc_frames) {
      // pc/RotationFunctions.py:428-508 (free SO(3) branch): one normalised Gaussian quaternion per (point, frame)
      HB_CHECK_ARENA("frames");
      if (n == 0) return SE3_OK;
      HB_TRY(se3_quat_frames(u_frames + 4 * (int64_t)F * u_off, n * F, ar.at<float>(c.frames), s));
      u_off += n;
      return se3_pack_records(ar.at<float>(c.pts), ar.at<float>(c.frames), n, F, ar.at<float>(c.rec), s);
    }
    const int64_t o_knn = ar.take((size_t)n * d->knn_k * 4);
    const size_t wsb = se3_knn_workspace_bytes(n);
    const int64_t o_ws = ar.take(wsb);
    HB_CHECK_ARENA("frames");
    if (n == 0) return SE3_OK;
    HB_TRY(knn_query_impl(ar.at<float>(c.pts), ar.at<int32_t>(c.batch), n, d->knn_k, ar.at<char>(o_ws), wsb,
                          ar.at<int32_t>(o_knn), d->n_batches, max_seg, s, ar.at<float>(o_rawmin[raw_slot]),
                          ar.at<float>(o_rawmax[raw_slot])));
    HB_TRY(pca_frames_select_pack(ar.at<float>(c.pts), ar.at<int32_t>(o_knn), n, d->knn_k, d->fixed_axis, u_frames + u_off, F,
                                  ar.at<float>(c.frames), ar.at<float>(c.rec), s));
    u_off += n;
    return SE3_OK;
  };
  // cloud `dst` = grid-average pooling of (p, b) over grid g
  auto pool_cloud = [&](const float* p, const int32_t* b, int64_t n, const se3_hier_cloud& g, se3_hier_cloud& dst) -> int {
    dst.n = g.m;
    if (level_fused) return SE3_OK;  // pooled by the level kernel
    dst.pts = ar.take((size_t)g.m * 12);
    dst.batch = ar.take((size_t)g.m * 4);
    HB_CHECK_ARENA("pooled cloud");
    HB_TRY(se3_segment_pool_f32(p, n, 3, ar.at<int64_t>(g.sorted_ids), ar.at<int32_t>(g.cell_ends), g.m, 0,
                                ar.at<float>(dst.pts), stream));
    HB_TRY(se3_segment_first_i32(b, ar.at<int64_t>(g.sorted_ids), ar.at<int32_t>(g.cell_ends), g.m,
                                 ar.at<int32_t>(dst.batch), stream));
    return SE3_OK;
  };

  // ---- side streams (see Lanes): fork from the caller's stream
  float h_rad[SE3_HIER_MAX_NEIGH * 4];
  for (int i = 0; i < d->n_neigh; ++i) {
    SE3_CHECK_ARG(d->neigh_src[i] >= 0 && d->neigh_src[i] < n_clouds && d->neigh_dst[i] >= 0 && d->neigh_dst[i] < n_clouds,
                  "neighbourhood refers to a cloud that does not exist");
    SE3_CHECK_ARG(d->neigh_radius[i] > 0.0f, "radius must be positive");
    h_rad[4 * i] = h_rad[4 * i + 1] = h_rad[4 * i + 2] = d->neigh_radius[i];
    h_rad[4 * i + 3] = 0.0f;
  }
  float* d_rad = ar.at<float>(o_rad);
  if (d->n_neigh > 0) {
    // pageable source: the copy is staged before the call returns, so the stack buffer is safe
    SE3_CUDA(cudaMemcpyAsync(d_rad, h_rad, sizeof(float) * 4 * d->n_neigh, cudaMemcpyHostToDevice, st));
  }
  // everything after the fork runs inside `body`, so that EVERY exit path (errors, arena retries) joins the side
  // streams before the caller can free or reuse the arena
  auto body = [&]() -> int {
  SE3_CUDA(cudaEventRecord(ln->fork, st));
  for (int k = 0; k < kSide; ++k) SE3_CUDA(cudaStreamWaitEvent(ln->s[k], ln->fork, 0));
  int side_rr = 0;
  // joins every side stream back into the caller's stream (before host reads and before returning)
  auto join_all = [&]() -> int {
    for (int k = 0; k < kSide; ++k) {
      SE3_CUDA(cudaEventRecord(ln->join[k], ln->s[k]));
      SE3_CUDA(cudaStreamWaitEvent(st, ln->join[k], 0));
    }
    return SE3_OK;
  };
  // ---- neighbourhoods: one sorted source structure per (source cloud, radius) and one count pass per
  // neighbourhood, issued on the side streams as soon as the clouds they touch exist (the large neighbourhoods
  // of the fine levels then run underneath the rest of the pooling chain); ONE blocking read of the edge totals
  // at the end, then fills + transposed rows
  struct Source {
    int cloud;
    float radius;
    int64_t ws, mn, mx, nc;
    size_t ws_bytes;
    bool issued;
  } sources[SE3_HIER_MAX_NEIGH];
  int n_sources = 0, src_of[SE3_HIER_MAX_NEIGH], rad_slot[SE3_HIER_MAX_NEIGH];
  for (int i = 0; i < d->n_neigh; ++i) {
    int f = -1;
    for (int j = 0; j < n_sources; ++j)
      if (sources[j].cloud == d->neigh_src[i] && sources[j].radius == d->neigh_radius[i]) f = j;
    if (f < 0) {
      f = n_sources++;
      sources[f].cloud = d->neigh_src[i];
      sources[f].radius = d->neigh_radius[i];
      sources[f].issued = false;
      rad_slot[f] = i;
    }
    src_of[i] = f;
  }
  bool cloud_done[SE3_HIER_MAX_CLOUDS + 1] = {false}, nb_issued[SE3_HIER_MAX_NEIGH] = {false};
  int batch_rr = 0;
  int64_t wd_off[SE3_HIER_MAX_NEIGH];
  size_t wd_bytes[SE3_HIER_MAX_NEIGH];
  auto issue_ready = [&]() -> int {
    for (int j = 0; j < n_sources; ++j) {
      Source& so = sources[j];
      if (so.issued || !cloud_done[so.cloud]) continue;
      const se3_hier_cloud& s = out->clouds[so.cloud];
      so.ws_bytes = se3_ball_query_src_workspace_bytes(s.n, 1);
      so.ws = ar.take(so.ws_bytes);
      so.mn = ar.take(bb);
      so.mx = ar.take(bb);
      so.nc = ar.take(16);
      HB_CHECK_ARENA("ball-query source");
      cudaStream_t sj = ln->s[j % kSide];
      se3_stream_t sjs = reinterpret_cast<se3_stream_t>(sj);
      SE3_CUDA(cudaStreamWaitEvent(sj, ln->cloud[so.cloud], 0));
      if (s.n > 0 && seg_build_possible(d->n_batches, cloud_seg[so.cloud])) {
        HB_TRY(ball_query_prepare_fused(ar.at<float>(s.pts), ar.at<int32_t>(s.batch), s.n,
                                        ar.at<float>(o_rawmin[1 + so.cloud]), ar.at<float>(o_rawmax[1 + so.cloud]),
                                        so.radius, ar.at<float>(so.mn), ar.at<float>(so.mx), ar.at<int32_t>(so.nc),
                                        ar.at<char>(so.ws), so.ws_bytes, d->n_batches, cloud_seg[so.cloud], sjs));
      } else {
        HB_TRY(se3_grid_extents(ar.at<float>(o_rawmin[1 + so.cloud]), ar.at<float>(o_rawmax[1 + so.cloud]), d->n_batches,
                                so.radius, -1e-6f, ar.at<float>(so.mn), ar.at<float>(so.mx), ar.at<int32_t>(so.nc), sjs));
        HB_TRY(ball_query_prepare_impl(ar.at<float>(s.pts), ar.at<int32_t>(s.batch), s.n, 1, ar.at<float>(so.mn),
                                       ar.at<int32_t>(so.nc), d_rad + 4 * rad_slot[j], ar.at<char>(so.ws), so.ws_bytes,
                                       key_bits_for(so.radius), d->n_batches, cloud_seg[so.cloud], sjs));
      }
      SE3_CUDA(cudaEventRecord(ln->src[j], sj));
      so.issued = true;
    }
    // every neighbourhood that has become possible goes into ONE batched count launch (+ one batched scan launch)
    int ready[SE3_HIER_MAX_NEIGH], n_ready = 0;
    for (int i = 0; i < d->n_neigh; ++i)
      if (!nb_issued[i] && sources[src_of[i]].issued && cloud_done[d->neigh_dst[i]]) ready[n_ready++] = i;
    for (int r0 = 0; r0 < n_ready; r0 += kBqBatch) {
      const int nb_n = n_ready - r0 < kBqBatch ? n_ready - r0 : kBqBatch;
      BqBatchItem items[kBqBatch];
      // the transposed row arrays of a batch are taken back to back: one memset clears them all
      const int64_t t_block = (int64_t)ar.off;
      for (int k = 0; k < nb_n; ++k) out->neigh[ready[r0 + k]].t_row_ends = ar.take((size_t)out->clouds[d->neigh_src[ready[r0 + k]]].n * 4);
      const size_t t_block_bytes = ar.off - (size_t)t_block;
      cudaStream_t sb = ln->s[batch_rr++ % kSide];
      for (int k = 0; k < nb_n; ++k) {
        const int i = ready[r0 + k];
        const Source& so = sources[src_of[i]];
        const se3_hier_cloud& s = out->clouds[d->neigh_src[i]];
        const se3_hier_cloud& t = out->clouds[d->neigh_dst[i]];
        se3_hier_neigh& nb = out->neigh[i];
        nb.row_ends = ar.take((size_t)t.n * 4);
        wd_bytes[i] = se3_ball_query_dst_workspace_bytes(t.n);
        wd_off[i] = ar.take(wd_bytes[i]);
        HB_CHECK_ARENA("ball-query workspace");
        SE3_CUDA(cudaStreamWaitEvent(sb, ln->src[src_of[i]], 0));
        SE3_CUDA(cudaStreamWaitEvent(sb, ln->cloud[d->neigh_dst[i]], 0));
        BqBatchItem& q = items[k];
        memset(&q, 0, sizeof(q));
        q.pts_dst = ar.at<float>(t.pts); q.batch_dst = ar.at<int32_t>(t.batch);
        q.n_src = s.n; q.n_dst = t.n;
        q.min_pt = ar.at<float>(so.mn); q.num_cells = ar.at<int32_t>(so.nc); q.radius = d_rad + 4 * i;
        q.ws_src = ar.at<char>(so.ws); q.ws_src_bytes = so.ws_bytes;
        q.ws_dst = ar.at<char>(wd_off[i]); q.ws_dst_bytes = wd_bytes[i];
        q.row_ends = ar.at<int32_t>(nb.row_ends); q.t_row = ar.at<int32_t>(nb.t_row_ends);
        q.total_out = d_cnt + 1 + i;
        nb_issued[i] = true;
      }
      HB_CHECK_ARENA("ball-query workspace");
      HB_TRY(bq_count_transposed_batch(items, nb_n, ar.at<char>(t_block), t_block_bytes, reinterpret_cast<se3_stream_t>(sb)));
    }
    return SE3_OK;
  };
  // cloud c is complete on the caller's stream (points, batch ids): raw box + the event the side streams wait on
  auto cloud_mark = [&](int c, int raw_slot, bool have_box) -> int {
    if (!have_box) HB_TRY(cloud_bbox(out->clouds[c], raw_slot, stream));
    SE3_CUDA(cudaEventRecord(ln->cloud[c], st));
    return SE3_OK;
  };
  // ... its frames go to a side stream, and every source / count that has become possible is issued
  auto cloud_side = [&](int c) -> int {
    cudaStream_t s = ln->s[side_rr++ % kSide];
    SE3_CUDA(cudaStreamWaitEvent(s, ln->cloud[c], 0));
    HB_TRY(build_frames(out->clouds[c], cloud_seg[c], 1 + c, reinterpret_cast<se3_stream_t>(s)));
    SE3_CUDA(cudaEventRecord(ln->frames[c], s));
    cloud_done[c] = true;
    return issue_ready();
  };

  // ---- level 0 and the output cloud share the raw cloud's init_cell grid
  out->raw.n = d->n;
  HB_TRY(grid_launch(pts, batch_ids, d->n, 0, d->init_cell, out->raw, seg_raw, out->clouds[0], 1));
  HB_TRY(grid_finish(d->n, out->raw));
  cloud_seg[0] = seg_next;
  if (d->out_cloud) cloud_seg[d->n_pool + 1] = seg_next;  // one point per raw cell, like level 0
  HB_TRY(pool_cloud(pts, batch_ids, d->n, out->raw, out->clouds[0]));
  HB_TRY(cloud_mark(0, 1, level_fused));
  // ---- pooled levels.  Order per level: launch the grid of cloud l, THEN issue cloud l's side work (the host is the
  // bottleneck of this build: its launches now overlap the grid chain on the GPU), then block on the grid size.
  for (int l = 0; l <= d->n_pool; ++l) {
    se3_hier_cloud& src = out->clouds[l];
    if (l < d->n_pool) {
      SE3_CHECK_ARG(d->cells[l] > 0.0f, "cell sizes must be positive");
      HB_TRY(grid_launch(ar.at<float>(src.pts), ar.at<int32_t>(src.batch), src.n, 1 + l, d->cells[l], src, cloud_seg[l],
                         out->clouds[l + 1], 2 + l));
    }
    HB_TRY(cloud_side(l));
    if (l == 0 && d->out_cloud) {
      // only needs the raw grid: picked early, so its kNN / PCA (the largest frame job) overlaps the pooling chain.
      // Its frame variates are the last n of u_frames (the levels consume at most (n_pool + 1) * n).
      se3_hier_cloud& oc = out->clouds[d->n_pool + 1];
      oc.n = out->raw.m;
      oc.pts = ar.take((size_t)oc.n * 12);
      oc.batch = ar.take((size_t)oc.n * 4);
      out->out_picked = ar.take((size_t)oc.n * 8);
      HB_CHECK_ARENA("output cloud");
      HB_TRY(se3_segment_pick(pts, batch_ids, ar.at<int64_t>(out->raw.sorted_ids), ar.at<int32_t>(out->raw.cell_ends),
                              oc.n, u_cells, ar.at<float>(oc.pts), ar.at<int32_t>(oc.batch),
                              ar.at<int64_t>(out->out_picked), stream));
      const int64_t u_keep = u_off;
      u_off = (int64_t)(d->n_pool + 1) * d->n;
      HB_TRY(cloud_mark(d->n_pool + 1, d->n_pool + 2, false));
      HB_TRY(cloud_side(d->n_pool + 1));
      u_off = u_keep;
    }
    if (l < d->n_pool) {
      HB_TRY(grid_finish(src.n, src));
      cloud_seg[l + 1] = seg_next;
      HB_TRY(pool_cloud(ar.at<float>(src.pts), ar.at<int32_t>(src.batch), src.n, src, out->clouds[l + 1]));
      HB_TRY(cloud_mark(l + 1, 2 + l, level_fused));
    }
  }

  stamp("pooling chain done");
  HB_TRY(issue_ready());
  stamp("counts issued");
  HB_TRY(join_all());
  if (d->n_neigh > 0) {
    SE3_CUDA(cudaMemcpyAsync(host, d_cnt + 1, sizeof(int64_t) * d->n_neigh, cudaMemcpyDeviceToHost, st));
    SE3_CUDA(cudaStreamSynchronize(st));
  }
  stamp("edge totals read");
  for (int i0 = 0; i0 < d->n_neigh; i0 += kBqBatch) {
    const int nb_n = d->n_neigh - i0 < kBqBatch ? d->n_neigh - i0 : kBqBatch;
    BqBatchItem items[kBqBatch];
    for (int k = 0; k < nb_n; ++k) {
      const int i = i0 + k;
      const Source& so = sources[src_of[i]];
      const se3_hier_cloud& s = out->clouds[d->neigh_src[i]];
      const se3_hier_cloud& t = out->clouds[d->neigh_dst[i]];
      se3_hier_neigh& nb = out->neigh[i];
      nb.e = host[i];
      nb.col_src = ar.take((size_t)nb.e * 4);
      nb.edge_dst = ar.take((size_t)nb.e * 4);
      nb.t_edge = ar.take((size_t)nb.e * 4);
      nb.t_dst = ar.take((size_t)nb.e * 4);
      HB_CHECK_ARENA("neighbourhood");
      BqBatchItem& q = items[k];
      memset(&q, 0, sizeof(q));
      q.pts_dst = ar.at<float>(t.pts);
      q.n_src = s.n; q.n_dst = t.n; q.radius = d_rad + 4 * i;
      q.ws_src = ar.at<char>(so.ws); q.ws_src_bytes = so.ws_bytes;
      q.ws_dst = ar.at<char>(wd_off[i]); q.ws_dst_bytes = wd_bytes[i];
      q.row_ends = ar.at<int32_t>(nb.row_ends); q.t_row = ar.at<int32_t>(nb.t_row_ends);
      q.n_edges = nb.e;
      q.col_src = ar.at<int32_t>(nb.col_src); q.edge_dst = ar.at<int32_t>(nb.edge_dst);
      q.t_edge = ar.at<int32_t>(nb.t_edge); q.t_dst = ar.at<int32_t>(nb.t_dst);
    }
    // everything issued so far has completed (the blocking read above): all fills in one launch on the caller's
    // stream (blockIdx.y = neighbourhood), all row orderings in a second one
    HB_TRY(bq_fill_transposed_batch(items, nb_n, stream));
  }
    return SE3_OK;
  };
  const int rc = body();
  bool joined = true;
  for (int k = 0; k < kSide; ++k) {
    joined = joined && cudaEventRecord(ln->join[k], ln->s[k]) == cudaSuccess;
    joined = joined && cudaStreamWaitEvent(st, ln->join[k], 0) == cudaSuccess;
  }
  if (!joined) {
    for (int k = 0; k < kSide; ++k) cudaStreamSynchronize(ln->s[k]);
  }
  stamp("fills issued, joined");
  if (rc != SE3_OK) return rc;
  out->arena_used = (int64_t)ar.off;
  return SE3_OK;
}
