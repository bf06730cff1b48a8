// Interface of the fused tcgen05 convolution kernel (conv_fused.cu).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace se3 {

struct FusedArgs {
  const int* row_ends;          // [n_rows] inclusive CSR row ends
  const int* nbr;               // [n_edges] gathered point per CSR entry
  const float* rec_row;         // [n_rows * f_row, 12] (point, frame) records of the row side
  const float* rec_g;           // [n_g * f_g, 12] records of the gathered side
  int f_g;                      // frames per gathered point
  const __nv_bfloat16* feat;    // [n_g * f_g, cs] bf16 feature rows of the gathered side
  int cs;                       // row stride of feat in elements (multiple of 8)
  int c;                        // gathered channels
  const float* w9;              // proj_axes_ [9,32]
  const float* bias;            // proj_biases_ [32]
  float norm;                   // norm_neigh_dist_
  int act;
  float out_scale;
  const unsigned char* w3img;   // pre-swizzled projection weights (launch_w3_image), nullptr = no projection
  int co;                       // output channels of the projection
  float* out;                   // [n_rows * f_row, co] fp32 (projection mode)
  __nv_bfloat16* t_save;        // optional [n_rows * f_row, 32 * CP] bf16 copy of the aggregation tile, (k,c) order
  int64_t n_rows;
  int64_t n_edges;
};

// gathered channels -> padded channel count of the kernel instantiation (0 = not supported)
inline int fused_cp(int c) { return c <= 8 ? 0 : (c <= 16 ? 16 : (c <= 32 ? 32 : (c <= 64 ? 64 : 0))); }
int fused_mode();
void fused_set_mode(int m);
bool fused_supported(int c_gathered, int co, int f_row, int f_g, bool project);
size_t fused_w3_bytes(int c_gathered, int co);
// plain: row-major [co][32 * CP] instead of the swizzled k-block image
int launch_w3_image(const float* w, int c_in, int c_out, bool tr, bool plain, __nv_bfloat16* img, cudaStream_t st);
int launch_conv_fused(const FusedArgs& a, int f_row, bool tr, cudaStream_t st);

}  // namespace se3
