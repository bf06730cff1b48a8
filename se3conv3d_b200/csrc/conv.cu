// Top-level dispatch of the fused PNEConvLayerRotEquiv forward / backward (C ABI).
//
// precision 0 (fp32, CUDA cores):
//   fwd : T = aggregate(x)            [R, Cin*K]   (geometry + basis fused, saved for backward)
//         y = s * T . W                [R, Cout]
//   bwd : dT = s * dy . W^T            [R, Cin*K]
//         dW = s * T^T . dy            (deterministic split-K)
//         d(proj_axes, proj_biases)    by output row, per-warp partials + ordered reduce
//         U  = aggregate^T(dy)         [N*F_in, Cout*K]  over the transposed CSR (no atomics)
//         dx = s * U . Wp^T            [N*F_in, Cin]
// precision 1 (bf16 tensor cores) lives in conv_tc.cu.
#include "conv_simt.cuh"

namespace se3 {
// conv_tc.cu
size_t conv_tc_fwd_workspace_bytes(const se3_conv_desc* d);
size_t conv_tc_bwd_workspace_bytes(const se3_conv_desc* d);
size_t conv_tc_saved_bytes(const se3_conv_desc* d);
size_t conv_tc_weight_cache_bytes(const se3_conv_desc* d);
int conv_tc_fwd(const se3_conv_desc* d, const float* x, float* y, void* saved, void* ws, size_t ws_bytes,
                cudaStream_t st);
int conv_tc_bwd(const se3_conv_desc* d, const float* x, const float* dy, const void* saved, float* dx, float* dW,
                float* dA, float* dB, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace se3

using namespace se3;

static int check_desc(const se3_conv_desc* d) {
  SE3_CHECK_ARG(d != nullptr, "null descriptor");
  SE3_CHECK_ARG(d->k == 32, "only K = 32 basis functions is supported (all shipped models use 32)");
  SE3_CHECK_ARG(d->f_in >= 1 && d->f_in <= 4 && d->f_out >= 1 && d->f_out <= 4, "frames per point must be 1..4");
  SE3_CHECK_ARG(d->c_in >= 1 && d->c_out >= 1, "bad channel counts");
  SE3_CHECK_ARG(d->n_in >= 0 && d->n_out >= 0 && d->n_edges >= 0, "bad sizes");
  SE3_CHECK_ARG(d->act >= 0 && d->act <= 3, "unknown activation");
  SE3_CHECK_ARG(d->precision == 0 || d->precision == 1, "unknown precision");
  // shape limits of the tensor-core path are reported when the byte counts are queried, not in the middle of a call
  SE3_CHECK_ARG(d->precision == 0 || d->c_out % 8 == 0, "precision 1 needs c_out to be a multiple of 8");
  SE3_CHECK_ARG(d->precision == 0 || (d->rec_in && d->rec_out), "precision 1 needs the packed gather records (se3_pack_records)");
  return SE3_OK;
}

static int splits_for(int64_t m, int64_t n, int64_t k) {
  const int64_t tiles = ((m + 63) / 64) * ((n + 63) / 64);
  int64_t s = (4ll * num_sms() + tiles - 1) / tiles;
  if (s > 32) s = 32;
  if (s * 256 > k) s = k / 256;
  if (s < 1) s = 1;
  return (int)s;
}

extern "C" size_t se3_conv_saved_bytes(const se3_conv_desc* d) {
  if (check_desc(d) != SE3_OK) return 0;
  if (d->precision == 1) return conv_tc_saved_bytes(d);
  return align_up((size_t)d->n_out * d->f_out * d->c_in * d->k * sizeof(float)) + 256;
}

extern "C" size_t se3_conv_weight_cache_bytes(const se3_conv_desc* d) {
  if (check_desc(d) != SE3_OK || d->precision != 1) return 0;
  return conv_tc_weight_cache_bytes(d);
}

extern "C" size_t se3_conv_fwd_workspace_bytes(const se3_conv_desc* d) {
  if (check_desc(d) != SE3_OK) return 0;
  if (d->precision == 1) return conv_tc_fwd_workspace_bytes(d);
  return 256;
}

extern "C" size_t se3_conv_bwd_workspace_bytes(const se3_conv_desc* d) {
  if (check_desc(d) != SE3_OK) return 0;
  if (d->precision == 1) return conv_tc_bwd_workspace_bytes(d);
  const int64_t R = d->n_out * d->f_out, ck = (int64_t)d->c_in * d->k;
  size_t b = 0;
  b += align_up((size_t)R * ck * 4);                                       // dT
  b += align_up((size_t)d->n_in * d->f_in * d->c_out * d->k * 4);           // U
  b += align_up((size_t)ck * d->c_out * 4);                                 // Wp
  b += align_up((size_t)splits_for(ck, d->c_out, R) * ck * d->c_out * 4);   // dW partials
  b += align_up((size_t)edge_grad_partials(d->n_out) * 320 * 4);            // basis-gradient partials
  return b + 256;
}

extern "C" int se3_conv_fwd(const se3_conv_desc* d, const float* x, float* y, void* saved, void* workspace,
                            size_t workspace_bytes, se3_stream_t stream) {
  if (int rc = check_desc(d)) return rc;
  cudaStream_t st = as_stream(stream);
  const int64_t R = d->n_out * d->f_out;
  if (R == 0) return SE3_OK;
  SE3_CHECK_ARG(x && y, "null tensor");
  if (d->precision == 1) return conv_tc_fwd(d, x, y, saved, workspace, workspace_bytes, st);
  SE3_CHECK_ARG(saved != nullptr, "precision 0 needs the saved buffer (se3_conv_saved_bytes)");
  float* T = reinterpret_cast<float*>(saved);
  AggArgs a;
  a.row_ends = d->row_ends; a.nbr = d->col_src;
  a.pts_row = d->pts_out; a.frm_row = d->frames_out; a.f_row = d->f_out;
  a.pts_g = d->pts_in; a.frm_g = d->frames_in; a.f_g = d->f_in;
  a.feat = x; a.c = d->c_in; a.w9 = d->proj_axes; a.bias = d->proj_biases;
  a.norm = d->norm_neigh_dist; a.act = d->act; a.out = T; a.n_rows = d->n_out;
  if (int rc = launch_aggregate_f32(a, false, st)) return rc;
  const int64_t ck = (int64_t)d->c_in * d->k;
  return launch_sgemm(false, false, R, d->c_out, ck, d->out_scale, T, ck, d->conv_weights, d->c_out, y, d->c_out, 1,
                      nullptr, st);
}

extern "C" int se3_conv_bwd(const se3_conv_desc* d, const float* x, const float* dy, const void* saved, float* dx,
                            float* d_conv_weights, float* d_proj_axes, float* d_proj_biases, void* workspace,
                            size_t workspace_bytes, se3_stream_t stream) {
  if (int rc = check_desc(d)) return rc;
  cudaStream_t st = as_stream(stream);
  const int64_t R = d->n_out * d->f_out, ck = (int64_t)d->c_in * d->k;
  const int64_t Nf = d->n_in * d->f_in;
  if (R == 0 || d->n_edges == 0) {
    if (dx && Nf) SE3_CUDA(cudaMemsetAsync(dx, 0, Nf * d->c_in * sizeof(float), st));
    if (d_conv_weights) SE3_CUDA(cudaMemsetAsync(d_conv_weights, 0, ck * d->c_out * sizeof(float), st));
    if (d_proj_axes) SE3_CUDA(cudaMemsetAsync(d_proj_axes, 0, 9 * d->k * sizeof(float), st));
    if (d_proj_biases) SE3_CUDA(cudaMemsetAsync(d_proj_biases, 0, d->k * sizeof(float), st));
    return SE3_OK;
  }
  SE3_CHECK_ARG(x && dy && workspace, "null tensor");
  if (d->precision == 1)
    return conv_tc_bwd(d, x, dy, saved, dx, d_conv_weights, d_proj_axes, d_proj_biases, workspace, workspace_bytes, st);
  SE3_CHECK_ARG(saved != nullptr, "precision 0 backward needs the forward's saved buffer");
  SE3_CHECK_ARG(!dx || (d->t_row_ends && d->t_edge && d->t_dst), "dx needs the transposed CSR");
  const float* T = reinterpret_cast<const float*>(saved);
  Arena ar(workspace, workspace_bytes);
  float* dT = ar.take<float>(R * ck);
  float* U = ar.take<float>(Nf * d->c_out * d->k);
  float* Wp = ar.take<float>(ck * d->c_out);
  const int splits = splits_for(ck, d->c_out, R);
  float* dWp = ar.take<float>((size_t)splits * ck * d->c_out);
  const int n_part = edge_grad_partials(d->n_out);
  float* eg = ar.take<float>((size_t)n_part * 320);
  if (!ar.ok()) {
    set_error("se3_conv_bwd: workspace too small");
    return SE3_EWORKSPACE;
  }
  if (d_conv_weights) {
    if (int rc = launch_sgemm(true, false, ck, d->c_out, R, d->out_scale, T, ck, dy, d->c_out, d_conv_weights,
                              d->c_out, splits, dWp, st))
      return rc;
  }
  if (d_proj_axes || d_proj_biases) {
    if (int rc = launch_sgemm(false, true, R, ck, d->c_out, d->out_scale, dy, d->c_out, d->conv_weights, d->c_out, dT,
                              ck, 1, nullptr, st))
      return rc;
    EdgeGradArgs g;
    g.row_ends = d->row_ends; g.col_src = d->col_src;
    g.pts_out = d->pts_out; g.frm_out = d->frames_out; g.f_out = d->f_out;
    g.pts_in = d->pts_in; g.frm_in = d->frames_in; g.f_in = d->f_in;
    g.x = x; g.c = d->c_in; g.w9 = d->proj_axes; g.bias = d->proj_biases;
    g.norm = d->norm_neigh_dist; g.act = d->act; g.dT = dT; g.n_out = d->n_out;
    g.partials = eg; g.n_partials = n_part;
    if (int rc = launch_edge_grad_f32(g, d_proj_axes, d_proj_biases, st)) return rc;
  }
  if (dx) {
    AggArgs a;
    a.row_ends = d->t_row_ends; a.nbr = d->t_dst;
    a.pts_row = d->pts_in; a.frm_row = d->frames_in; a.f_row = d->f_in;
    a.pts_g = d->pts_out; a.frm_g = d->frames_out; a.f_g = d->f_out;
    a.feat = dy; a.c = d->c_out; a.w9 = d->proj_axes; a.bias = d->proj_biases;
    a.norm = d->norm_neigh_dist; a.act = d->act; a.out = U; a.n_rows = d->n_in;
    if (int rc = launch_aggregate_f32(a, true, st)) return rc;
    if (int rc = launch_permute_w(d->conv_weights, d->c_in, d->k, d->c_out, Wp, st)) return rc;
    const int64_t ok = (int64_t)d->c_out * d->k;
    if (int rc = launch_sgemm(false, true, Nf, d->c_in, ok, d->out_scale, U, ok, Wp, ok, dx, d->c_in, 1, nullptr, st))
      return rc;
  }
  return SE3_OK;
}
