// Micro-benchmark: tcgen05.ld throughput per SM as a function of the number of reading warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../lvt_b200/csrc/common.cuh"
void lvt_set_error(const char*, ...) {}
bool lvt_pdl_enabled() { return false; }

__global__ void k(int iters, long long* out, float* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r0[32], r1[32], r2[32], r3[32];
    const uint32_t c = ((i + warp) * 128) & 511;
    tmem_ld_32x32(base + c, r0);
    tmem_ld_32x32(base + ((c + 32) & 511), r1);
    tmem_ld_32x32(base + ((c + 64) & 511), r2);
    tmem_ld_32x32(base + ((c + 96) & 511), r3);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += __uint_as_float(r0[j] ^ r1[j] ^ r2[j] ^ r3[j]);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tptr, 512); }
}

int main() {
  long long* d; float* s;
  cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 16}) {
    k<<<1, warps * 32>>>(iters, d, s);
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * 4 * 4096 * warps;
    printf("warps %2d: %lld cycles, %.1f B/clk/SM, %.1f clk per 32x32b.x32 load per warp\n", warps, h, bytes / h,
           (double)h / (iters * 4));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
