// Micro-benchmark: cost of back-to-back tcgen05.mma instructions issued by one thread (clocks per instruction, measured
// from the first issue to the completion of the commit) for the operand layouts / shapes the fused kernel uses.
#include <cstdio>
#include <cuda_runtime.h>
#include "../se3conv3d_b200/csrc/umma.cuh"
using namespace se3::umma;
namespace se3 { void set_error(const char*, ...) {} void count_launch(int) {} bool pdl_enabled() { return false; } }

template <int MODE>
__global__ void __launch_bounds__(128, 1) k_bench(long long* out, int reps) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_addr(smem_dyn) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024; i += 128) smem_dyn[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_addr(&bar), 1); mbar_init_fence(); }
  if (threadIdx.x < 32) tmem_alloc(smem_addr(&slot), 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    fence_proxy_async();
    uint64_t ad, bd; uint32_t id;
    if (MODE == 0) { ad = desc_kmajor_sw128(base); bd = desc_kmajor_sw128(base + 16384); id = idesc(FMT_BF16, FMT_BF16, 128, 32, false, false); }
    if (MODE == 1) { ad = desc_mnmajor_sw128(base, 16384); bd = desc_mnmajor_sw128(base + 32768, 8192); id = idesc(FMT_BF16, FMT_BF16, 128, 32, true, true); }
    if (MODE == 2) { ad = desc_kmajor_sw128(base); bd = desc_kmajor_sw128(base + 16384); id = idesc(FMT_TF32, FMT_TF32, 128, 32, false, false); }
    if (MODE == 3) { ad = desc_kmajor_sw128(base); bd = desc_kmajor_sw128(base + 16384); id = idesc(FMT_BF16, FMT_BF16, 128, 16, false, false); }
    if (MODE == 4) { ad = desc_kmajor_sw128(base); bd = desc_kmajor_sw128(base + 16384); id = idesc(FMT_BF16, FMT_BF16, 128, 256, false, false); }
    if (MODE == 5) { ad = desc_mnmajor_sw128(base, 16384); bd = desc_mnmajor_sw128(base + 32768, 8192); id = idesc(FMT_BF16, FMT_BF16, 128, 64, true, true); }
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        if (MODE == 2) mma_tf32(tmem, ad, bd, id, 1u); else mma_f16(tmem, ad, bd, id, 1u);
      }
      const long long t1 = clock64();
      commit(smem_addr(&bar));
      mbar_wait(smem_addr(&bar), (uint32_t)(rep & 1));
      const long long t2 = clock64();
      out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

template <int MODE> void run(const char* name, long long* d) {
  const int reps = 256;
  cudaFuncSetAttribute(k_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_bench<MODE><<<1, 128, 100 * 1024>>>(d, reps);
  long long h[6];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s issue %.1f clk/mma, issue+complete %.1f clk/mma   (%s)\n", name, h[4] / (double)reps, h[5] / (double)reps,
         cudaGetErrorString(e));
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<0>("bf16 K-major  M128 N32  K16", d);
  run<1>("bf16 MN-major M128 N32  K16", d);
  run<5>("bf16 MN-major M128 N64  K16", d);
  run<2>("tf32 K-major  M128 N32  K8", d);
  run<3>("bf16 K-major  M128 N16  K16", d);
  run<4>("bf16 K-major  M128 N256 K16", d);
  return 0;
}
