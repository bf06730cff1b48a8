// fp32 CUDA-core ("exactness mode") building blocks of the fused PNEConvLayerRotEquiv path, and
// helpers shared with the tensor-core path.
#pragma once
#include "common.cuh"

namespace se3 {

// Arguments of the fused geometry -> basis -> aggregate kernels.
//   forward  (TR=false): rows = output points (frame a), gathered = input points (frame b)
//   transposed (TR=true): rows = input points (frame b), gathered = output points (frame a)
// In both cases the basis is h[e,a,b,:] = act(g . W9 + bias) with
//   d = (p_in - p_out) * norm,  g = [ d^T R_out,a ; rows 0-1 of R_out,a^T R_in,b ]
// (layers/PNEConvLayerRotEquiv.py:68-90, pc/RotationFunctions.py:637-665, 549-600, 236-252).
struct AggArgs {
  const int* row_ends;   // [n_rows] inclusive
  const int* nbr;        // [E] gathered point per CSR entry
  const float* pts_row;
  const float* frm_row;
  int f_row;
  const float* pts_g;
  const float* frm_g;
  int f_g;
  const float* feat;     // [n_g * f_g, c]
  int c;
  const float* w9;       // [9,32]
  const float* bias;     // [32]
  float norm;
  int act;
  float* out;            // [n_rows * f_row, c, 32]
  int64_t n_rows;
};

int launch_aggregate_f32(const AggArgs& a, bool transposed, cudaStream_t st);

// C[M,N] = alpha * op(A) * op(B); AT: A stored [K,M]; BT: B stored [N,K].  Deterministic split-K
// through `partials` (size splits*M*N floats) when splits > 1.
int launch_sgemm(bool at, bool bt, int64_t m, int64_t n, int64_t k, float alpha, const float* a, int64_t lda,
                 const float* b, int64_t ldb, float* c, int64_t ldc, int splits, float* partials, cudaStream_t st);

struct EdgeGradArgs {
  const int* row_ends;
  const int* col_src;
  const float* pts_out;
  const float* frm_out;
  int f_out;
  const float* pts_in;
  const float* frm_in;
  int f_in;
  const float* x;    // [n_in*f_in, c]
  int c;
  const float* w9;
  const float* bias;
  float norm;
  int act;
  const float* dT;   // [n_out*f_out, c, 32]
  int64_t n_out;
  float* partials;   // [n_partials, 10, 32]
  int n_partials;
};
int edge_grad_partials(int64_t n_out);
int launch_edge_grad_f32(const EdgeGradArgs& a, float* d_axes, float* d_bias, cudaStream_t st);

// wp[c][o][k] = w[c][k][o]
int launch_permute_w(const float* w, int c_in, int k, int c_out, float* wp, cudaStream_t st);

}  // namespace se3
