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
void profile_begin(int id, cudaStream_t st, void** handle);
void profile_end(void* handle, cudaStream_t st);
struct ProfScope {
  void* h;
  cudaStream_t st;
  ProfScope(int id, cudaStream_t s) : st(s) { profile_begin(id, s, &h); }
  ~ProfScope() { profile_end(h, st); }
};

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
