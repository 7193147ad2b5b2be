// gvk_optim.cu — fused global-norm clip + Adam on the flat trainable-parameter buffer (see include/gvk.h).
// Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step of reference src/train.py:315-319 (a few hundred foreach
// launches over 304 tensors) with two launches over one contiguous buffer.
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, size_t n, float scale, float* __restrict__ partials) {
  __shared__ float red[256];
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t begin = blockIdx.x * per, end = min(n, begin + per);
  float s = 0.f;
  for (size_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const float v = g[i] * scale;
    s = fmaf(v, v, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                                                          const float* __restrict__ partials, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                                                          float eps, float wd, float bc1, float bc2_sqrt, float* __restrict__ norm_out) {
  __shared__ float s_coef;
  if (threadIdx.x == 0) {
    float coef = 1.f;
    if (partials) {
      float tot = 0.f;
      for (int i = 0; i < GVK_SUMSQ_PARTIALS; ++i) tot += partials[i];  // fixed order: identical on every block / rank
      const float norm = sqrtf(tot);
      if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
      if (norm_out && blockIdx.x == 0) *norm_out = norm;
    }
    s_coef = coef * grad_scale;
  }
  __syncthreads();
  const float coef = s_coef;
  const float step_size = lr / bc1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * coef);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

// Same update with the learning rate and the 1-based step read from device memory (CUDA-graph replay: the values change between replays
// of the same kernel node); the bias corrections are formed on the device with the host formula's operations (powf, sqrtf).
__global__ void __launch_bounds__(256) clip_adam_dyn_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                                                              const float* __restrict__ partials, float max_norm, float grad_scale, const float* __restrict__ lr_dev,
                                                              float beta1, float beta2, float eps, float wd, const long long* __restrict__ step_dev,
                                                              float* __restrict__ norm_out) {
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    float coef = 1.f;
    if (partials) {
      float tot = 0.f;
      for (int i = 0; i < GVK_SUMSQ_PARTIALS; ++i) tot += partials[i];
      const float norm = sqrtf(tot);
      if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
      if (norm_out && blockIdx.x == 0) *norm_out = norm;
    }
    s_coef = coef * grad_scale;
    const float step = static_cast<float>(*step_dev);
    s_step_size = *lr_dev / (1.f - powf(beta1, step));
    s_bc2_sqrt = sqrtf(1.f - powf(beta2, step));
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * coef);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

int clip_adam_dyn(float* param, const float* grad, float* m, float* v, size_t n, const float* partials, float max_norm, float grad_scale, const float* lr_dev,
                  float beta1, float beta2, float eps, float wd, const long long* step_dev, float* norm_out, cudaStream_t stream) {
  GVK_CHECK_ARG(param && grad && m && v && n > 0 && lr_dev && step_dev, "gvk_clip_adam_dyn: bad argument");
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 4);
  clip_adam_dyn_kernel<<<grid, 256, 0, stream>>>(param, grad, m, v, n, partials, max_norm, grad_scale, lr_dev, beta1, beta2, eps, wd, step_dev, norm_out);
  GVK_CHECK_LAUNCH("clip_adam_dyn");
  return GVK_OK;
}

int grad_sumsq(const float* grad, size_t n, float grad_scale, float* partials, cudaStream_t stream) {
  GVK_CHECK_ARG(grad && partials && n > 0, "gvk_grad_sumsq: bad argument");
  grad_sumsq_kernel<<<GVK_SUMSQ_PARTIALS, 256, 0, stream>>>(grad, n, grad_scale, partials);
  GVK_CHECK_LAUNCH("grad_sumsq");
  return GVK_OK;
}

int clip_adam(float* param, const float* grad, float* m, float* v, size_t n, const float* partials, float max_norm, float grad_scale, float lr, float beta1,
              float beta2, float eps, float wd, int step, float* norm_out, cudaStream_t stream) {
  GVK_CHECK_ARG(param && grad && m && v && n > 0 && step >= 1, "gvk_clip_adam: bad argument");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 4);
  clip_adam_kernel<<<grid, 256, 0, stream>>>(param, grad, m, v, n, partials, max_norm, grad_scale, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, norm_out);
  GVK_CHECK_LAUNCH("clip_adam");
  return GVK_OK;
}

}  // namespace gvk
