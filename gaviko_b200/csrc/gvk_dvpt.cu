// gvk_dvpt.cu — kernels of the DVPT side path (reference model/dvpt.py:25-47, share_MLP):
//   * QuickGELU applied BEFORE the rank-r down-projection (elementwise forward / backward-with-residual),
//   * the prompt -> image-token cross attention in the rank-r latent (queries = the prompt latents themselves, no query projection,
//     scale d_model^-0.5, keys = values = all N token latents), forward and backward,
//   * the scalar prompt_gate: scaled copies of the up-projection parameters and the gradients of (W_u, b_u, gate) from the gradients of the
//     scaled parameters.
// All fp32 SIMT: the latent is 20 wide, the work per volume is ~4 MFLOP, the kernels are bound by staging the token latents (80 KB / volume).
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {

// -------------------------------------------------------------------------------------------------
// elementwise
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) quickgelu_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    y[i] = make_float4(quick_gelu(v.x), quick_gelu(v.y), quick_gelu(v.z), quick_gelu(v.w));
  }
}
int quickgelu_fwd(const float* x, float* y, size_t n, cudaStream_t stream) {
  GVK_CHECK_ARG(x && y && n > 0 && n % 4 == 0, "gvk_quickgelu_fwd: bad argument (n %% 4 == 0)");
  GVK_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "gvk_quickgelu_fwd: 16-byte alignment");
  const int grid = (int)std::min<size_t>((n / 4 + 255) / 256, (size_t)sm_count() * 16);
  quickgelu_fwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n / 4);
  GVK_CHECK_LAUNCH("quickgelu_fwd");
  return GVK_OK;
}

__global__ void __launch_bounds__(256) quickgelu_bwd_add_kernel(const float4* __restrict__ dy, const float4* __restrict__ pre, const float4* __restrict__ res,
                                                                  float4* __restrict__ y, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 d = dy[i], p = pre[i];
    float4 o = make_float4(d.x * quick_gelu_grad(p.x), d.y * quick_gelu_grad(p.y), d.z * quick_gelu_grad(p.z), d.w * quick_gelu_grad(p.w));
    if (res) {
      const float4 r = res[i];
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    y[i] = o;
  }
}
int quickgelu_bwd_add(const float* dy, const float* pre, const float* res, float* y, size_t n, cudaStream_t stream) {
  GVK_CHECK_ARG(dy && pre && y && n > 0 && n % 4 == 0, "gvk_quickgelu_bwd_add: bad argument (n %% 4 == 0)");
  GVK_CHECK_ARG(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
                "gvk_quickgelu_bwd_add: 16-byte alignment");
  const int grid = (int)std::min<size_t>((n / 4 + 255) / 256, (size_t)sm_count() * 16);
  quickgelu_bwd_add_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(pre), reinterpret_cast<const float4*>(res),
                                                     reinterpret_cast<float4*>(y), n / 4);
  GVK_CHECK_LAUNCH("quickgelu_bwd_add");
  return GVK_OK;
}

// -------------------------------------------------------------------------------------------------
// prompt -> token cross attention in the latent (R = 20)
// -------------------------------------------------------------------------------------------------
constexpr int kXaR = 20;
constexpr int kXaThreads = 512;

__device__ __forceinline__ float dot20(const float* __restrict__ row, const float (&q)[kXaR]) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kXaR; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + c);
    s = fmaf(q[c], t.x, s); s = fmaf(q[c + 1], t.y, s); s = fmaf(q[c + 2], t.z, s); s = fmaf(q[c + 3], t.w, s);
  }
  return s;
}
__device__ __forceinline__ void axpy20(float a, const float* __restrict__ row, float (&acc)[kXaR]) {
#pragma unroll
  for (int c = 0; c < kXaR; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + c);
    acc[c] = fmaf(a, t.x, acc[c]); acc[c + 1] = fmaf(a, t.y, acc[c + 1]); acc[c + 2] = fmaf(a, t.z, acc[c + 2]); acc[c + 3] = fmaf(a, t.w, acc[c + 3]);
  }
}
__device__ __forceinline__ void load20(const float* __restrict__ row, float (&v)[kXaR]) {
#pragma unroll
  for (int c = 0; c < kXaR; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + c);
    v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
  }
}
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int nfloats) {
  for (int i = threadIdx.x * 4; i < nfloats; i += blockDim.x * 4) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(src + i);
}

// One CTA per volume; the N token latents are staged once in shared memory; a warp owns one prompt at a time (online softmax over the lanes'
// key slices, merged with shuffles).  The attention output replaces the prompt's latent row in place; the query it was is saved in pl.
__global__ void __launch_bounds__(kXaThreads) latent_xattn_fwd_kernel(gvk_latent_xattn_fwd_params p) {
  extern __shared__ __align__(16) float xa_smem[];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.T - p.P - 1;
  float* zb = p.z + (size_t)b * p.T * kXaR;
  stage_rows(xa_smem, zb + (size_t)(p.P + 1) * kXaR, N * kXaR);
  __syncthreads();
  for (int pp = warp; pp < p.P; pp += kXaThreads / 32) {
    float q[kXaR];
    load20(zb + (size_t)pp * kXaR, q);
    const float q_lane = lane < kXaR ? zb[(size_t)pp * kXaR + lane] : 0.f;     // the same row, one element per lane, for the saved copy
    float m = -INFINITY, l = 0.f, acc[kXaR];
#pragma unroll
    for (int c = 0; c < kXaR; ++c) acc[c] = 0.f;
    for (int t = lane; t < N; t += 32) {
      const float* row = xa_smem + t * kXaR;
      const float s = dot20(row, q) * p.scale;
      if (s > m) {
        const float corr = __expf(m - s);
        l *= corr;
#pragma unroll
        for (int c = 0; c < kXaR; ++c) acc[c] *= corr;
        m = s;
      }
      const float pr = __expf(s - m);
      l += pr;
      axpy20(pr, row, acc);
    }
    const float M = warp_max(m);
    const float corr = (m == -INFINITY) ? 0.f : __expf(m - M);
    l = warp_sum(l * corr);
    const float inv_l = 1.0f / l;
    float out = 0.f;
#pragma unroll
    for (int c = 0; c < kXaR; ++c) {
      const float v = warp_sum(acc[c] * corr) * inv_l;
      if (lane == c) out = v;
    }
    __syncwarp();
    if (lane < kXaR) {
      p.pl[((size_t)b * p.P + pp) * kXaR + lane] = q_lane;
      zb[(size_t)pp * kXaR + lane] = out;
    }
    if (lane == 0) p.lse[(size_t)b * p.P + pp] = M + __logf(l);
  }
}

// Backward, one CTA per volume, two phases over the staged token latents:
//   A (a warp per prompt, lanes over keys): dq[p] = scale * sum_n ds[p, n] k[n],   ds = P (dP - delta), dP = dctx . k, delta = dctx . ctx
//   B (a thread per key, loop over prompts): dk[n] = dcomb[n] + sum_p (scale * ds[p, n] q[p] + P[p, n] dctx[p])
// dz enters as d(combined latent) and leaves as d(latent before the attention), in place; the cls row passes through.
__global__ void __launch_bounds__(kXaThreads) latent_xattn_bwd_kernel(gvk_latent_xattn_bwd_params p) {
  extern __shared__ __align__(16) float xa_smem[];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.T - p.P - 1, P = p.P;
  float* s_k = xa_smem;                  // [N][R]
  float* s_q = s_k + (size_t)N * kXaR;   // [P][R]
  float* s_dc = s_q + P * kXaR;          // [P][R]  d ctx
  float* s_lse = s_dc + P * kXaR;        // [P]
  float* s_delta = s_lse + P;            // [P]
  const float* zb = p.z + (size_t)b * p.T * kXaR;
  float* dzb = p.dz + (size_t)b * p.T * kXaR;
  stage_rows(s_k, zb + (size_t)(P + 1) * kXaR, N * kXaR);
  stage_rows(s_q, p.pl + (size_t)b * P * kXaR, P * kXaR);
  stage_rows(s_dc, dzb, P * kXaR);
  __syncthreads();
  for (int pp = threadIdx.x; pp < P; pp += kXaThreads) {
    float d = 0.f;
    for (int c = 0; c < kXaR; ++c) d = fmaf(s_dc[pp * kXaR + c], zb[(size_t)pp * kXaR + c], d);    // rows < P of z hold ctx
    s_delta[pp] = d;
    s_lse[pp] = p.lse[(size_t)b * P + pp];
  }
  __syncthreads();
  // ---- phase A
  for (int pp = warp; pp < P; pp += kXaThreads / 32) {
    float q[kXaR], dc[kXaR], acc[kXaR];
    load20(s_q + pp * kXaR, q);
    load20(s_dc + pp * kXaR, dc);
#pragma unroll
    for (int c = 0; c < kXaR; ++c) acc[c] = 0.f;
    const float lse = s_lse[pp], delta = s_delta[pp];
    for (int t = lane; t < N; t += 32) {
      const float* row = s_k + t * kXaR;
      const float pr = __expf(dot20(row, q) * p.scale - lse);
      const float ds = pr * (dot20(row, dc) - delta);
      axpy20(ds, row, acc);
    }
    float out = 0.f;
#pragma unroll
    for (int c = 0; c < kXaR; ++c) {
      const float v = warp_sum(acc[c]);
      if (lane == c) out = v;
    }
    if (lane < kXaR) dzb[(size_t)pp * kXaR + lane] = out * p.scale;
  }
  // ---- phase B
  for (int t = threadIdx.x; t < N; t += kXaThreads) {
    float k[kXaR], dk[kXaR];
    load20(s_k + t * kXaR, k);
    float* drow = dzb + (size_t)(P + 1 + t) * kXaR;
    load20(drow, dk);
    for (int pp = 0; pp < P; ++pp) {
      const float* q = s_q + pp * kXaR;
      const float* dc = s_dc + pp * kXaR;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < kXaR; ++c) {
        s = fmaf(q[c], k[c], s);
        dp = fmaf(dc[c], k[c], dp);
      }
      const float pr = __expf(s * p.scale - s_lse[pp]);
      const float ds = pr * (dp - s_delta[pp]) * p.scale;
#pragma unroll
      for (int c = 0; c < kXaR; ++c) dk[c] = fmaf(ds, q[c], fmaf(pr, dc[c], dk[c]));
    }
#pragma unroll
    for (int c = 0; c < kXaR; c += 4) *reinterpret_cast<float4*>(drow + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
  }
}

static int xattn_check(const void* z, int B, int T, int P, int r, size_t smem, const char* who) {
  GVK_CHECK_ARG(z && B > 0 && P > 0 && T > P + 1, "%s: bad shape B=%d T=%d P=%d", who, B, T, P);
  GVK_CHECK_ARG(r == kXaR, "%s: latent width %d (DVPT hard-codes 20, model/dvpt.py:27)", who, r);
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0, "%s: 16-byte alignment", who);
  GVK_CHECK_ARG(smem <= 220 * 1024, "%s: %d token latents do not fit shared memory", who, T - P - 1);
  return GVK_OK;
}

int latent_xattn_fwd(const gvk_latent_xattn_fwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->pl && p->lse, "gvk_latent_xattn_fwd: null pointer");
  const size_t smem = (size_t)(p->T - p->P - 1) * kXaR * sizeof(float);
  int st = xattn_check(p->z, p->B, p->T, p->P, p->r, smem, "gvk_latent_xattn_fwd");
  if (st != GVK_OK) return st;
  st = cuda_status(cudaFuncSetAttribute(latent_xattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "latent_xattn_fwd smem");
  if (st != GVK_OK) return st;
  latent_xattn_fwd_kernel<<<p->B, kXaThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("latent_xattn_fwd");
  return GVK_OK;
}

int latent_xattn_bwd(const gvk_latent_xattn_bwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->pl && p->lse && p->dz, "gvk_latent_xattn_bwd: null pointer");
  const size_t smem = ((size_t)(p->T - p->P - 1) * kXaR + 2 * (size_t)p->P * kXaR + 2 * p->P) * sizeof(float);
  int st = xattn_check(p->z, p->B, p->T, p->P, p->r, smem, "gvk_latent_xattn_bwd");
  if (st != GVK_OK) return st;
  GVK_CHECK_ARG(((size_t)p->P * kXaR) % 4 == 0 && ((reinterpret_cast<uintptr_t>(p->dz) | reinterpret_cast<uintptr_t>(p->pl)) & 15) == 0,
                "gvk_latent_xattn_bwd: P * r must be a multiple of 4, 16-byte alignment");
  st = cuda_status(cudaFuncSetAttribute(latent_xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "latent_xattn_bwd smem");
  if (st != GVK_OK) return st;
  latent_xattn_bwd_kernel<<<p->B, kXaThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("latent_xattn_bwd");
  return GVK_OK;
}

// -------------------------------------------------------------------------------------------------
// scalar gate on a parameter block:  y = gate * x   and   dx += gate * dy,  dgate += <x, dy>   (deterministic, one CTA)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_scale_kernel(const float* __restrict__ x, const float* __restrict__ gate, float* __restrict__ y, size_t n) {
  const float g = gate[0];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = g * x[i];
}
int gate_scale(const float* x, const float* gate, float* y, size_t n, cudaStream_t stream) {
  GVK_CHECK_ARG(x && gate && y && n > 0, "gvk_gate_scale: bad argument");
  gate_scale_kernel<<<(int)std::min<size_t>((n + 255) / 256, 1024), 256, 0, stream>>>(x, gate, y, n);
  GVK_CHECK_LAUNCH("gate_scale");
  return GVK_OK;
}

__global__ void __launch_bounds__(1024) gate_grads_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gate, float* __restrict__ dx,
                                                            float* __restrict__ dgate, size_t n) {
  __shared__ float s_part[32];
  const float g = gate[0];
  float acc = 0.f;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = dy[i];
    acc = fmaf(x[i], d, acc);
    dx[i] += g * d;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    const float v = warp_sum(s_part[threadIdx.x]);
    if (threadIdx.x == 0) dgate[0] += v;
  }
}
int gate_grads(const float* x, const float* dy, const float* gate, float* dx, float* dgate, size_t n, cudaStream_t stream) {
  GVK_CHECK_ARG(x && dy && gate && dx && dgate && n > 0, "gvk_gate_grads: bad argument");
  gate_grads_kernel<<<1, 1024, 0, stream>>>(x, dy, gate, dx, dgate, n);
  GVK_CHECK_LAUNCH("gate_grads");
  return GVK_OK;
}

}  // namespace gvk
