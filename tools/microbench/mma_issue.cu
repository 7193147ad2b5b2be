// Microbenchmark (B200): cycles per tcgen05.mma (kind::f16, M = 128, K = 16) issued back to back by one thread, as a function of N,
// with the A operand in shared memory (SS) or tensor memory (TS).  Operands are whatever is in smem / TMEM: only timing matters.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../gaviko_b200/csrc/gvk_common.cuh"
using namespace gvk;

template <int N, bool TS>
__global__ void __launch_bounds__(128) k_mma(int iters, long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) umma_bf16_ts(tmem, tmem + 256 + 8 * k, make_sw128_desc(b_addr + k * 32, 16, 1024), idesc, 1u);
        else umma_bf16(tmem, make_sw128_desc(a_addr + k * 32, 16, 1024), make_sw128_desc(b_addr + k * 32, 16, 1024), idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cyc[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
template <int N, bool TS>
void run(long long* cyc) {
  const int iters = 2000;
  cudaFuncSetAttribute(k_mma<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k_mma<N, TS><<<1, 128, 64 * 1024>>>(iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("M=128 N=%3d K=16 %s: %.1f clk per MMA (math floor %d clk)  %s\n", N, TS ? "TS" : "SS", (double)h / (iters * 4), 128 * N / 256, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long* cyc; cudaMalloc(&cyc, 8);
  run<64, false>(cyc); run<128, false>(cyc); run<256, false>(cyc);
  run<64, true>(cyc); run<128, true>(cyc);
  run<32, false>(cyc); run<16, false>(cyc);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
