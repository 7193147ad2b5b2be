// Microbenchmark (B200): issue rate of the SFU exponentials per SM as a function of the operand format and of the number of warps per
// sub-partition.  Question: does ex2.approx.ftz.bf16x2 (two exponentials per lane per instruction) run at the warp-instruction rate of
// ex2.approx.ftz.f32 (i.e. twice the exponentials per clock), or is it split into two MUFU passes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sfu_rate sfu_rate.cu && ./sfu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k_sfu(int iters, long long* cyc, float* sink) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {        // fp32 ex2: 4 independent chains
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) { // bf16x2 ex2
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b3));
    } else if (MODE == 2) { // f16x2 ex2
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b3));
    } else if (MODE == 3) { // fp32 tanh
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
    } else {                // bf16x2 tanh
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(b0));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(b1));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(b2));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(b3));
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(b0 ^ b1 ^ b2 ^ b3);
}

template <int MODE>
void run(const char* name, int lanes_per_instr, long long* cyc, float* sink) {
  const int iters = 4000;
  for (int warps : {4, 8, 16, 32}) {
    k_sfu<MODE><<<1, warps * 32>>>(iters, cyc, sink);
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double instr = (double)iters * 4 * warps;      // warp instructions on one SM
    printf("%-28s %2d warps: %6.2f clk per warp instruction per SM -> %5.1f exponentials / clk / SM   %s\n", name, warps, h / instr,
           instr * 32 * lanes_per_instr / h, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 8 * 64);
  float* sink; cudaMalloc(&sink, 4 * 64 * 1024);
  run<0>("ex2.approx.ftz.f32", 1, cyc, sink);
  run<1>("ex2.approx.ftz.bf16x2", 2, cyc, sink);
  run<2>("ex2.approx.f16x2", 2, cyc, sink);
  run<3>("tanh.approx.f32", 1, cyc, sink);
  run<4>("tanh.approx.bf16x2", 2, cyc, sink);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
