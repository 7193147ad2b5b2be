// Microbenchmarks (B200): tcgen05.ld / tcgen05.st throughput per SM, MUFU ex2 throughput, for sizing the attention kernels.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../gaviko_b200/csrc/gvk_common.cuh"
using namespace gvk;

__global__ void __launch_bounds__(512) k_ldtm(int iters, int nwarps, float* sink, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t t = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      float v[32];
      tmem_ld_32x32(t + ((i * 32) & 255) + (warp >> 2) * 256 % 256, v);
      tc_wait_ld();
      acc += v[0] + v[31];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
// 4 independent x32 loads in flight before the wait
__global__ void __launch_bounds__(512) k_ldtm4(int iters, int nwarps, float* sink, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t t = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      float v0[32], v1[32], v2[32], v3[32];
      tmem_ld_32x32(t + 0, v0); tmem_ld_32x32(t + 32, v1); tmem_ld_32x32(t + 64, v2); tmem_ld_32x32(t + 96, v3);
      tc_wait_ld();
      acc += v0[0] + v1[31] + v2[5] + v3[7];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
__global__ void __launch_bounds__(512) k_ex2(int iters, float* sink, long long* cyc) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fast_ex2(x[k]) - 1.0f;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456f) sink[0] = s;
}
int main() {
  float* sink; long long* cyc; cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
  long long h;
  const int iters = 4096;
  for (int nw : {1, 4, 8, 16}) {
    k_ldtm<<<1, 512>>>(iters, nw, sink, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDTM.x32 one-at-a-time, %2d warps: %.1f clk per load per warp, %.1f B/clk/SM\n", nw, (double)h / iters, (double)nw * iters * 4096 / h);
    k_ldtm4<<<1, 512>>>(iters, nw, sink, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDTM.x32 4 in flight,     %2d warps: %.1f clk per 4 loads per warp, %.1f B/clk/SM\n", nw, (double)h / iters, (double)nw * iters * 4 * 4096 / h);
  }
  for (int nt : {128, 256, 512}) {
    k_ex2<<<1, nt>>>(iters, sink, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ex2.approx %3d threads: %.2f ex2/clk/SM\n", nt, (double)nt * iters * 8 / h);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
