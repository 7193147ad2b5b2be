// gvk_input.cu — the step in front of the hot path: per-volume intensity rescale to [out_min, out_max]
// (torchio.RescaleIntensity(out_min_max=(0,1)) at reference train.py:53,57,61 / eval.py:31 / inference.py:30, run there on the CPU inside
// the DataLoader workers).  On the device it is two HBM-bound passes: per-volume min / max partials (deterministic, no atomics), then
// y = (x - min) / (max - min) * (out_max - out_min) + out_min with IEEE fp32 division, in the operation order of torchio so the fp32
// result is bit-identical to the CPU transform.  A constant volume (max == min) is copied through unchanged, as torchio does.
#include "gvk_common.cuh"

namespace gvk {

constexpr int kRescaleThreads = 256;

__global__ void __launch_bounds__(kRescaleThreads) rescale_minmax_kernel(const float* __restrict__ in, long long n, float* __restrict__ ws) {
  const int b = blockIdx.y, part = blockIdx.x, parts = gridDim.x;
  const float* x = in + (size_t)b * n;
  const long long per = ((n + parts - 1) / parts + 3) & ~3LL;
  const long long lo = part * per, hi = min(n, lo + per);
  float mn = INFINITY, mx = -INFINITY;
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (vec) {
    for (long long i = lo + 4LL * threadIdx.x; i + 3 < hi; i += 4LL * kRescaleThreads) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      mn = fminf(fminf(mn, v.x), fminf(fminf(v.y, v.z), v.w));
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
    }
    for (long long i = lo + ((hi - lo) & ~3LL) + threadIdx.x; i < hi; i += kRescaleThreads) {
      mn = fminf(mn, x[i]);
      mx = fmaxf(mx, x[i]);
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += kRescaleThreads) {
      mn = fminf(mn, x[i]);
      mx = fmaxf(mx, x[i]);
    }
  }
  __shared__ float smn[kRescaleThreads / 32], smx[kRescaleThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < kRescaleThreads / 32; ++w) {
      mn = fminf(mn, smn[w]);
      mx = fmaxf(mx, smx[w]);
    }
    ws[((size_t)b * parts + part) * 2] = mn;
    ws[((size_t)b * parts + part) * 2 + 1] = mx;
  }
}

template <typename T>
__global__ void __launch_bounds__(kRescaleThreads) rescale_apply_kernel(const float* __restrict__ in, long long n, const float* __restrict__ ws, int parts, float out_min,
                                                                         float out_max, T* __restrict__ out) {
  const int b = blockIdx.y;
  float mn = INFINITY, mx = -INFINITY;
  for (int k = 0; k < parts; ++k) {   // 64 partials, broadcast loads
    mn = fminf(mn, ws[((size_t)b * parts + k) * 2]);
    mx = fmaxf(mx, ws[((size_t)b * parts + k) * 2 + 1]);
  }
  const float range = mx - mn, out_range = out_max - out_min;
  const bool identity = range == 0.f;
  const float* x = in + (size_t)b * n;
  T* y = out + (size_t)b * n;
  auto f = [&](float v) { return identity ? v : __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v, mn), range), out_range), out_min); };
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & (4 * sizeof(T) - 1)) == 0;
  const long long stride = (long long)gridDim.x * kRescaleThreads;
  if (vec) {
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * kRescaleThreads + threadIdx.x; i < n4; i += stride) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      const float4 r = make_float4(f(v.x), f(v.y), f(v.z), f(v.w));
      if constexpr (sizeof(T) == 4) {
        reinterpret_cast<float4*>(y)[i] = r;
      } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
        reinterpret_cast<uint2*>(y)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * kRescaleThreads + threadIdx.x; i < n; i += stride) st_from_float<T>(y + i, f(x[i]));
  } else {
    for (long long i = (long long)blockIdx.x * kRescaleThreads + threadIdx.x; i < n; i += stride) st_from_float<T>(y + i, f(x[i]));
  }
}

int rescale_intensity(const gvk_rescale_intensity_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->in && p->out && p->ws, "gvk_rescale_intensity: null pointer");
  GVK_CHECK_ARG(p->B > 0 && p->n > 0, "gvk_rescale_intensity: bad shape B=%d n=%lld", p->B, p->n);
  GVK_CHECK_ARG(p->out_dtype == GVK_F32 || p->out_dtype == GVK_BF16, "gvk_rescale_intensity: bad out_dtype");
  GVK_CHECK_ARG(p->out_dtype == GVK_F32 || p->out != (const void*)p->in, "gvk_rescale_intensity: in-place needs an fp32 output");
  rescale_minmax_kernel<<<dim3(GVK_RESCALE_PARTS, p->B), kRescaleThreads, 0, stream>>>(p->in, p->n, p->ws);
  GVK_CHECK_LAUNCH("rescale_intensity (min / max)");
  const int ctas = (int)std::max<long long>(1, std::min<long long>((p->n / 4 + kRescaleThreads - 1) / kRescaleThreads, std::max(1, sm_count() * 8 / p->B)));
  if (p->out_dtype == GVK_F32)
    rescale_apply_kernel<float><<<dim3(ctas, p->B), kRescaleThreads, 0, stream>>>(p->in, p->n, p->ws, GVK_RESCALE_PARTS, p->out_min, p->out_max, reinterpret_cast<float*>(p->out));
  else
    rescale_apply_kernel<__nv_bfloat16><<<dim3(ctas, p->B), kRescaleThreads, 0, stream>>>(p->in, p->n, p->ws, GVK_RESCALE_PARTS, p->out_min, p->out_max,
                                                                                        reinterpret_cast<__nv_bfloat16*>(p->out));
  GVK_CHECK_LAUNCH("rescale_intensity (apply)");
  return GVK_OK;
}

}  // namespace gvk
