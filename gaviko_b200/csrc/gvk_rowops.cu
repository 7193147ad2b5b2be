// gvk_rowops.cu — one-warp-per-token-row kernels: LayerNorm fwd/bwd, rank-r down/up projections, rank-r weight gradients.
// All fp32 math; these are the HBM-bound side paths of GAViKO (LocalSelfAttention / Awakening_Prompt projections,
// model/gaviko.py:149-187,229-244) and the LayerNorms around the frozen GEMMs (model/vision_transformer.py:30,49).
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {

constexpr int kRowWarps = 8;
constexpr int kRowThreads = kRowWarps * 32;

// Replayable dropout multiplier for element index e (see include/gvk.h).
__device__ __forceinline__ float drop_mult(uint64_t seed, uint64_t e, float p, float inv_keep) {
  const uint64_t ctr = e >> 2;
  const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t sel = (uint32_t)(e & 3);
  const uint32_t v = sel == 0 ? r.x : sel == 1 ? r.y : sel == 2 ? r.z : r.w;
  return u32_to_unit(v) >= p ? inv_keep : 0.f;
}
// Two consecutive elements (e even) share one Philox call.
__device__ __forceinline__ float2 drop_mult2(uint64_t seed, uint64_t e, float p, float inv_keep) {
  const uint64_t ctr = e >> 2;
  const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const bool hi = (e & 2) != 0;
  const uint32_t a = hi ? r.z : r.x, b = hi ? r.w : r.y;
  return make_float2(u32_to_unit(a) >= p ? inv_keep : 0.f, u32_to_unit(b) >= p ? inv_keep : 0.f);
}

#define GVK_DISPATCH_NITER(dim, ...)                                             \
  switch ((dim) / 64) {                                                          \
    case 3: { constexpr int NITER = 3; __VA_ARGS__; break; }                     \
    case 6: { constexpr int NITER = 6; __VA_ARGS__; break; }                     \
    case 12: { constexpr int NITER = 12; __VA_ARGS__; break; }                   \
    case 16: { constexpr int NITER = 16; __VA_ARGS__; break; }                   \
    default:                                                                     \
      set_last_error("row kernels support dim in {192, 384, 768, 1024}, got %d", (int)(dim)); \
      return GVK_ERR_UNSUPPORTED;                                                \
  }

static inline int row_grid(int M, size_t smem_bytes = 0) {
  const int blocks = (M + kRowWarps - 1) / kRowWarps;
  // kernels that stage a weight panel in smem amortise the staging over many rows: one resident wave only
  int per_sm = 8;
  if (smem_bytes > 0) per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / (smem_bytes + 1024)));
  const int cap = sm_count() * per_sm;
  return blocks < cap ? blocks : cap;
}

#ifdef __CUDACC__
// Stage the rank-r weight panel as sw[j * dim + c] = w(j, c), reading global memory in its own linear order (coalesced).
__device__ __forceinline__ void stage_weight(float* sw, const float* __restrict__ w, int r, int dim, int w_sj, int w_sc) {
  const int total = r * dim;
  if (w_sc == 1) {  // [r, dim] row-major
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int j = idx / dim, c = idx - j * dim;
      sw[idx] = w[(size_t)j * w_sj + c];
    }
  } else {          // [dim, r] row-major (w_sj == 1): source-linear index is c * w_sc + j
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int c = idx / r, j = idx - c * r;
      sw[j * dim + c] = w[(size_t)c * w_sc + (size_t)j * w_sj];
    }
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// LayerNorm forward
// ------------------------------------------------------------------------------------------------
template <int NITER>
__global__ void __launch_bounds__(kRowThreads) layernorm_fwd_kernel(gvk_layernorm_fwd_params p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dim = NITER * 64;
  const float inv_dim = 1.0f / dim;
  for (int row = blockIdx.x * kRowWarps + warp; row < p.M; row += gridDim.x * kRowWarps) {
    float2 xv[NITER];
    const float* xr = p.x + (size_t)row * p.ldx;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      xv[i] = *reinterpret_cast<const float2*>(xr + lane * 2 + 64 * i);
      s += xv[i].x + xv[i].y;
    }
    const float mean = warp_sum(s) * inv_dim;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const float a = xv[i].x - mean, b = xv[i].y - mean;
      v += a * a + b * b;
    }
    const float rstd = rsqrtf(warp_sum(v) * inv_dim + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[row] = mean;
      if (p.rstd) p.rstd[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const int c = lane * 2 + 64 * i;
      const float2 g = *reinterpret_cast<const float2*>(p.gamma + c);
      const float2 b = *reinterpret_cast<const float2*>(p.beta + c);
      float y0 = (xv[i].x - mean) * rstd * g.x + b.x;
      float y1 = (xv[i].y - mean) * rstd * g.y + b.y;
      if (p.ssf_scale) {
        y0 = y0 * p.ssf_scale[c] + p.ssf_shift[c];
        y1 = y1 * p.ssf_scale[c + 1] + p.ssf_shift[c + 1];
      }
      if (p.y_dtype == GVK_F32) {
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.y) + (size_t)row * p.ldy + c) = make_float2(y0, y1);
      } else {
        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)row * p.ldy + c) = __floats2bfloat162_rn(y0, y1);
      }
    }
  }
}

int layernorm_fwd(const gvk_layernorm_fwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->x && p->y && p->gamma && p->beta, "gvk_layernorm_fwd: null pointer");
  GVK_CHECK_ARG(p->M > 0 && p->ldx % 2 == 0 && p->ldy % 2 == 0, "gvk_layernorm_fwd: bad shape");
  GVK_DISPATCH_NITER(p->dim, layernorm_fwd_kernel<NITER><<<row_grid(p->M), kRowThreads, 0, stream>>>(*p));
  GVK_CHECK_LAUNCH("layernorm_fwd");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// rank-r down projection (+ optional LN prologue, activation, chained second projection)
// ------------------------------------------------------------------------------------------------
template <int NITER>
__global__ void __launch_bounds__(kRowThreads) rowproj_down_kernel(gvk_rowproj_down_params p) {
  extern __shared__ float smem[];
  const int dim = NITER * 64;
  float* sw = smem;                 // [r][dim]
  float* sw2 = smem + p.r * dim;    // [r2][r]
  stage_weight(sw, p.w, p.r, dim, p.w_sj, p.w_sc);
  if (p.w2)
    for (int idx = threadIdx.x; idx < p.r2 * p.r; idx += blockDim.x) sw2[idx] = p.w2[idx];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_dim = 1.0f / dim;
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  for (int row = blockIdx.x * kRowWarps + warp; row < p.M; row += gridDim.x * kRowWarps) {
    float2 xv[NITER];
    const float* xr = p.x + (size_t)row * p.ldx;
#pragma unroll
    for (int i = 0; i < NITER; ++i) xv[i] = *reinterpret_cast<const float2*>(xr + lane * 2 + 64 * i);
    if (p.drop_p > 0.f) {
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const float2 m = drop_mult2(p.seed, p.offset + (uint64_t)row * dim + lane * 2 + 64 * i, p.drop_p, inv_keep);
        xv[i].x *= m.x;
        xv[i].y *= m.y;
      }
    }
    if (p.ln_gamma) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NITER; ++i) s += xv[i].x + xv[i].y;
      const float mean = warp_sum(s) * inv_dim;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const float a = xv[i].x - mean, b = xv[i].y - mean;
        v += a * a + b * b;
      }
      const float rstd = rsqrtf(warp_sum(v) * inv_dim + p.eps);
      if (lane == 0) {
        if (p.mean) p.mean[row] = mean;
        if (p.rstd) p.rstd[row] = rstd;
      }
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const int c = lane * 2 + 64 * i;
        const float2 g = *reinterpret_cast<const float2*>(p.ln_gamma + c);
        const float2 b = *reinterpret_cast<const float2*>(p.ln_beta + c);
        xv[i].x = (xv[i].x - mean) * rstd * g.x + b.x;
        xv[i].y = (xv[i].y - mean) * rstd * g.y + b.y;
      }
    }
    float zl = 0.f;  // lane j keeps z[j]
    for (int j = 0; j < p.r; ++j) {
      const float* wj = sw + j * dim + lane * 2;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const float2 w = *reinterpret_cast<const float2*>(wj + 64 * i);
        acc = fmaf(xv[i].x, w.x, acc);
        acc = fmaf(xv[i].y, w.y, acc);
      }
      acc = warp_sum(acc);
      if (lane == j) zl = acc;
    }
    if (lane < p.r) {
      float pre = zl + (p.bias ? p.bias[lane] : 0.f);
      if (p.pre) p.pre[(size_t)row * p.ldz + lane] = pre;
      if (p.act == GVK_ROWACT_QUICKGELU)
        pre = quick_gelu(pre);
      else if (p.act == GVK_ROWACT_RELU)
        pre = fmaxf(pre, 0.f);
      zl = pre;
      p.z[(size_t)row * p.ldz + lane] = zl;
    } else {
      zl = 0.f;
    }
    if (p.w2) {
      float o[3] = {0.f, 0.f, 0.f};
      for (int j = 0; j < p.r; ++j) {
        const float zj = __shfl_sync(0xffffffffu, zl, j);
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (lane + 32 * u < p.r2) o[u] = fmaf(zj, sw2[(lane + 32 * u) * p.r + j], o[u]);
      }
#pragma unroll
      for (int u = 0; u < 3; ++u)
        if (lane + 32 * u < p.r2) p.z2[(size_t)row * p.ldz2 + lane + 32 * u] = o[u];
    }
  }
}

int rowproj_down(const gvk_rowproj_down_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->x && p->w && p->z, "gvk_rowproj_down: null pointer");
  GVK_CHECK_ARG(p->r >= 1 && p->r <= 32 && p->r2 >= 0 && p->r2 <= 96, "gvk_rowproj_down: r=%d (<=32), r2=%d (<=96)", p->r, p->r2);
  GVK_CHECK_ARG(p->M > 0 && p->ldx % 2 == 0, "gvk_rowproj_down: bad shape");
  const size_t smem = ((size_t)p->r * p->dim + (size_t)p->r2 * p->r) * sizeof(float);
  GVK_DISPATCH_NITER(p->dim, {
    static size_t configured = 0;
    if (smem > configured) {
      int st = cuda_status(cudaFuncSetAttribute(rowproj_down_kernel<NITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "rowproj_down smem");
      if (st != GVK_OK) return st;
      configured = smem;
    }
    rowproj_down_kernel<NITER><<<row_grid(p->M, smem), kRowThreads, smem, stream>>>(*p);
  });
  GVK_CHECK_LAUNCH("rowproj_down");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// rank-r up projection + bias + dropout + residual
// ------------------------------------------------------------------------------------------------
template <int NITER>
__global__ void __launch_bounds__(kRowThreads) rowproj_up_kernel(gvk_rowproj_up_params p) {
  extern __shared__ float smem[];
  const int dim = NITER * 64;
  float* sw = smem;  // [r][dim]
  stage_weight(sw, p.w, p.r, dim, p.w_sj, p.w_sc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  for (int row = blockIdx.x * kRowWarps + warp; row < p.M; row += gridDim.x * kRowWarps) {
    const float cl = lane < p.r ? p.c[(size_t)row * p.ldc + lane] : 0.f;
    float2 acc[NITER];
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const int c = lane * 2 + 64 * i;
      acc[i] = p.bias ? *reinterpret_cast<const float2*>(p.bias + c) : make_float2(0.f, 0.f);
    }
    for (int j = 0; j < p.r; ++j) {
      const float cj = __shfl_sync(0xffffffffu, cl, j);
      const float* wj = sw + j * dim + lane * 2;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const float2 w = *reinterpret_cast<const float2*>(wj + 64 * i);
        acc[i].x = fmaf(cj, w.x, acc[i].x);
        acc[i].y = fmaf(cj, w.y, acc[i].y);
      }
    }
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const int c = lane * 2 + 64 * i;
      float2 v = acc[i];
      if (p.drop_p > 0.f) {
        const float2 m = drop_mult2(p.seed, p.offset + (uint64_t)row * dim + c, p.drop_p, inv_keep);
        v.x *= m.x;
        v.y *= m.y;
      }
      if (p.res) {
        const float2 r = *reinterpret_cast<const float2*>(p.res + (size_t)row * p.ld_res + c);
        v.x += r.x;
        v.y += r.y;
      }
      *reinterpret_cast<float2*>(p.out + (size_t)row * p.ld_out + c) = v;
      if (p.out_lp) *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.out_lp) + (size_t)row * p.ld_out_lp + c) = __floats2bfloat162_rn(v.x, v.y);
    }
  }
}

int rowproj_up(const gvk_rowproj_up_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->c && p->w && p->out, "gvk_rowproj_up: null pointer");
  GVK_CHECK_ARG(p->r >= 1 && p->r <= 32, "gvk_rowproj_up: r=%d must be in [1,32]", p->r);
  GVK_CHECK_ARG(p->M > 0 && p->ld_out % 2 == 0 && (!p->res || p->ld_res % 2 == 0), "gvk_rowproj_up: bad shape");
  const size_t smem = (size_t)p->r * p->dim * sizeof(float);
  GVK_DISPATCH_NITER(p->dim, {
    static size_t configured = 0;
    if (smem > configured) {
      int st = cuda_status(cudaFuncSetAttribute(rowproj_up_kernel<NITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "rowproj_up smem");
      if (st != GVK_OK) return st;
      configured = smem;
    }
    rowproj_up_kernel<NITER><<<row_grid(p->M, smem), kRowThreads, smem, stream>>>(*p);
  });
  GVK_CHECK_LAUNCH("rowproj_up");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// rank-r weight gradient  dw(j,c) += sum_m a[m,j] * f(x[m,c])
// ------------------------------------------------------------------------------------------------
constexpr int kWgRows = 8;  // rows staged per step

template <int NCOL, int R>
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(gvk_skinny_wgrad_params p, int rows_per_cta) {
  __shared__ float sa[kWgRows][R];
  __shared__ float smean[kWgRows], srstd[kWgRows];
  const int tid = threadIdx.x;
  const int m_begin = blockIdx.x * rows_per_cta;
  const int m_end = min(p.M, m_begin + rows_per_cta);
  float acc[NCOL][R];
  float xsum[NCOL];
#pragma unroll
  for (int i = 0; i < NCOL; ++i) {
    xsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;
  }
  float asum = 0.f;  // thread j < r
  float g[NCOL], b[NCOL];
#pragma unroll
  for (int i = 0; i < NCOL; ++i) {
    const int c = tid + 256 * i;
    g[i] = (p.ln_gamma && c < p.dim) ? p.ln_gamma[c] : 1.f;
    b[i] = (p.ln_gamma && c < p.dim) ? p.ln_beta[c] : 0.f;
  }
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  for (int m0 = m_begin; m0 < m_end; m0 += kWgRows) {
    __syncthreads();
    for (int idx = tid; idx < kWgRows * R; idx += 256) {
      const int rr = idx / R, j = idx - rr * R;
      const int m = m0 + rr;
      sa[rr][j] = (m < m_end && j < p.r) ? p.a[(size_t)m * p.lda + j] : 0.f;
    }
    if (tid < kWgRows) {
      const int m = m0 + tid;
      smean[tid] = (p.mean && m < m_end) ? p.mean[m] : 0.f;
      srstd[tid] = (p.rstd && m < m_end) ? p.rstd[m] : 1.f;
    }
    __syncthreads();
    if (tid < R) {
#pragma unroll
      for (int rr = 0; rr < kWgRows; ++rr) asum += sa[rr][tid];
    }
#pragma unroll
    for (int rr = 0; rr < kWgRows; ++rr) {
      const int m = m0 + rr;
      if (m >= m_end) break;
#pragma unroll
      for (int i = 0; i < NCOL; ++i) {
        const int c = tid + 256 * i;
        if (c < p.dim) {
          float x = p.x[(size_t)m * p.ldx + c];
          if (p.drop_p > 0.f) x *= drop_mult(p.seed, p.offset + (uint64_t)m * p.dim + c, p.drop_p, inv_keep);
          if (p.ln_gamma) x = (x - smean[rr]) * srstd[rr] * g[i] + b[i];
          xsum[i] += x;
#pragma unroll
          for (int j = 0; j < R; ++j) acc[i][j] = fmaf(sa[rr][j], x, acc[i][j]);
        }
      }
    }
  }
  // per-CTA partials (no atomics: same-address contention from ~300 CTAs serialises in L2); reduced by skinny_wgrad_reduce_kernel
  float* ws_dw = p.ws + (size_t)blockIdx.x * p.r * p.dim;
  float* ws_dx = p.ws + (size_t)gridDim.x * p.r * p.dim + (size_t)blockIdx.x * p.dim;
  float* ws_da = p.ws + (size_t)gridDim.x * (p.r + 1) * p.dim + (size_t)blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < NCOL; ++i) {
    const int c = tid + 256 * i;
    if (c < p.dim) {
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (j < p.r) ws_dw[(size_t)j * p.dim + c] = acc[i][j];
      ws_dx[c] = xsum[i];
    }
  }
  if (tid < 32) ws_da[tid] = tid < p.r ? asum : 0.f;
}

__global__ void __launch_bounds__(256) skinny_wgrad_reduce_kernel(gvk_skinny_wgrad_params p, int ncta) {
  const int rd = p.r * p.dim;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rd) {
    if (!p.dw) return;
    float s = 0.f;
    for (int k = 0; k < ncta; ++k) s += p.ws[(size_t)k * rd + idx];
    const int j = idx / p.dim, c = idx - j * p.dim;
    p.dw[(size_t)j * p.dw_sj + (size_t)c * p.dw_sc] += s;
  } else if (idx < rd + p.dim) {
    if (!p.dx_colsum) return;
    const int c = idx - rd;
    const float* base = p.ws + (size_t)ncta * rd;
    float s = 0.f;
    for (int k = 0; k < ncta; ++k) s += base[(size_t)k * p.dim + c];
    p.dx_colsum[c] += s;
  } else if (idx < rd + p.dim + p.r) {
    if (!p.da_colsum) return;
    const int j = idx - rd - p.dim;
    const float* base = p.ws + (size_t)ncta * (p.r + 1) * p.dim;
    float s = 0.f;
    for (int k = 0; k < ncta; ++k) s += base[(size_t)k * 32 + j];
    p.da_colsum[j] += s;
  }
}

static void skinny_wgrad_plan(int M, int* ctas, int* rows_per_cta) {
  const int want = std::max(1, std::min(sm_count() * 2, (M + 63) / 64));
  int rpc = (M + want - 1) / want;
  rpc = (rpc + kWgRows - 1) / kWgRows * kWgRows;
  *rows_per_cta = rpc;
  *ctas = (M + rpc - 1) / rpc;
}

size_t skinny_wgrad_ws_floats(int r, int dim, int M) {
  int ctas, rpc;
  skinny_wgrad_plan(M, &ctas, &rpc);
  return (size_t)ctas * ((size_t)(r + 1) * dim + 32);
}

template <int NCOL>
static int skinny_wgrad_launch(const gvk_skinny_wgrad_params* p, cudaStream_t stream) {
  int grid, rows_per_cta;
  skinny_wgrad_plan(p->M, &grid, &rows_per_cta);
  if (p->r <= 8)
    skinny_wgrad_kernel<NCOL, 8><<<grid, 256, 0, stream>>>(*p, rows_per_cta);
  else if (p->r <= 20)
    skinny_wgrad_kernel<NCOL, 20><<<grid, 256, 0, stream>>>(*p, rows_per_cta);
  else
    skinny_wgrad_kernel<NCOL, 32><<<grid, 256, 0, stream>>>(*p, rows_per_cta);
  GVK_CHECK_LAUNCH("skinny_wgrad");
  const int total = p->r * p->dim + p->dim + p->r;
  skinny_wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(*p, grid);
  GVK_CHECK_LAUNCH("skinny_wgrad_reduce");
  return GVK_OK;
}

int skinny_wgrad(const gvk_skinny_wgrad_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->a && p->x, "gvk_skinny_wgrad: null pointer");
  GVK_CHECK_ARG(p->r >= 1 && p->r <= 32 && p->dim >= 1 && p->dim <= 1024 && p->M > 0, "gvk_skinny_wgrad: r=%d dim=%d M=%d", p->r, p->dim, p->M);
  GVK_CHECK_ARG(!p->ln_gamma || (p->ln_beta && p->mean && p->rstd), "gvk_skinny_wgrad: LN recompute needs beta, mean, rstd");
  GVK_CHECK_ARG(p->ws && p->ws_floats >= skinny_wgrad_ws_floats(p->r, p->dim, p->M), "gvk_skinny_wgrad: workspace of %zu floats required (gvk_skinny_wgrad_ws_floats)",
                skinny_wgrad_ws_floats(p->r, p->dim, p->M));
  switch ((p->dim + 255) / 256) {
    case 1: return skinny_wgrad_launch<1>(p, stream);
    case 2: return skinny_wgrad_launch<2>(p, stream);
    case 3: return skinny_wgrad_launch<3>(p, stream);
    default: return skinny_wgrad_launch<4>(p, stream);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward
// ------------------------------------------------------------------------------------------------
template <int NITER, bool PGRAD>
__global__ void __launch_bounds__(kRowThreads) layernorm_bwd_kernel(gvk_layernorm_bwd_params p) {
  extern __shared__ float smem[];
  const int dim = NITER * 64;
  float* sw = smem;  // [r][dim] when dy is given in rank-r form
  if (p.dz) stage_weight(sw, p.w, p.r, dim, p.w_sj, p.w_sc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_dim = 1.0f / dim;
  constexpr int NG = PGRAD ? NITER : 1;
  float2 dg[NG], db[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) dg[i] = db[i] = make_float2(0.f, 0.f);
  float2 gam[NITER];
#pragma unroll
  for (int i = 0; i < NITER; ++i) gam[i] = *reinterpret_cast<const float2*>(p.gamma + lane * 2 + 64 * i);

  for (int row = blockIdx.x * kRowWarps + warp; row < p.M; row += gridDim.x * kRowWarps) {
    float2 dy[NITER];
    if (p.dz) {
      const float zl = lane < p.r ? p.dz[(size_t)row * p.ld_dz + lane] : 0.f;
#pragma unroll
      for (int i = 0; i < NITER; ++i) dy[i] = make_float2(0.f, 0.f);
      for (int j = 0; j < p.r; ++j) {
        const float zj = __shfl_sync(0xffffffffu, zl, j);
        const float* wj = sw + j * dim + lane * 2;
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          const float2 w = *reinterpret_cast<const float2*>(wj + 64 * i);
          dy[i].x = fmaf(zj, w.x, dy[i].x);
          dy[i].y = fmaf(zj, w.y, dy[i].y);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NITER; ++i) dy[i] = *reinterpret_cast<const float2*>(p.dy + (size_t)row * p.ld_dy + lane * 2 + 64 * i);
    }
    const float mean = p.mean[row], rstd = p.rstd[row];
    float2 xh[NITER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const float2 x = *reinterpret_cast<const float2*>(p.x + (size_t)row * p.ldx + lane * 2 + 64 * i);
      xh[i] = make_float2((x.x - mean) * rstd, (x.y - mean) * rstd);
      if constexpr (PGRAD) {
        dg[i].x += dy[i].x * xh[i].x;
        dg[i].y += dy[i].y * xh[i].y;
        db[i].x += dy[i].x;
        db[i].y += dy[i].y;
      }
      dy[i].x *= gam[i].x;  // g = dy * gamma
      dy[i].y *= gam[i].y;
      s1 += dy[i].x + dy[i].y;
      s2 += dy[i].x * xh[i].x + dy[i].y * xh[i].y;
    }
    const float m1 = warp_sum(s1) * inv_dim, m2 = warp_sum(s2) * inv_dim;
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const int c = lane * 2 + 64 * i;
      float2 dx = make_float2(rstd * (dy[i].x - m1 - xh[i].x * m2), rstd * (dy[i].y - m1 - xh[i].y * m2));
      if (p.dres) {
        const float2 r = *reinterpret_cast<const float2*>(p.dres + (size_t)row * p.ld_dres + c);
        dx.x += r.x;
        dx.y += r.y;
      }
      *reinterpret_cast<float2*>(p.dx + (size_t)row * p.ld_dx + c) = dx;
      if (p.dx_lp) *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.dx_lp) + (size_t)row * p.ld_dx_lp + c) = __floats2bfloat162_rn(dx.x, dx.y);
    }
  }
  if constexpr (PGRAD) {
    // cross-warp reduction through smem (reusing the weight staging area is unsafe: use a dedicated tail region)
    float* red = smem + (p.dz ? p.r * dim : 0);  // [kRowWarps][2*dim]
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const int c = lane * 2 + 64 * i;
      red[warp * 2 * dim + c] = dg[i].x;
      red[warp * 2 * dim + c + 1] = dg[i].y;
      red[warp * 2 * dim + dim + c] = db[i].x;
      red[warp * 2 * dim + dim + c + 1] = db[i].y;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * dim; idx += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) s += red[w * 2 * dim + idx];
      if (idx < dim) {
        if (p.dgamma) atomicAdd(p.dgamma + idx, s);
      } else {
        if (p.dbeta) atomicAdd(p.dbeta + idx - dim, s);
      }
    }
  }
}

int layernorm_bwd(const gvk_layernorm_bwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->x && p->gamma && p->mean && p->rstd && p->dx, "gvk_layernorm_bwd: null pointer");
  GVK_CHECK_ARG((p->dy != nullptr) != (p->dz != nullptr), "gvk_layernorm_bwd: exactly one of dy / dz must be given");
  GVK_CHECK_ARG(!p->dz || (p->w && p->r >= 1 && p->r <= 32), "gvk_layernorm_bwd: rank-r form needs w and 1 <= r <= 32");
  const bool red = p->dgamma || p->dbeta;
  const size_t smem = ((p->dz ? (size_t)p->r * p->dim : 0) + (red ? (size_t)kRowWarps * 2 * p->dim : 0)) * sizeof(float);
  // With parameter gradients every CTA ends with 2*dim atomics: keep the grid modest.
  int grid = row_grid(p->M, p->dz ? smem : 0);
  if (red) grid = std::min(grid, sm_count() * 2);
  GVK_DISPATCH_NITER(p->dim, {
    static size_t configured[2] = {0, 0};
    if (smem > configured[red]) {
      int st = red ? cuda_status(cudaFuncSetAttribute(layernorm_bwd_kernel<NITER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "layernorm_bwd smem")
                   : cuda_status(cudaFuncSetAttribute(layernorm_bwd_kernel<NITER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "layernorm_bwd smem");
      if (st != GVK_OK) return st;
      configured[red] = smem;
    }
    if (red)
      layernorm_bwd_kernel<NITER, true><<<grid, kRowThreads, smem, stream>>>(*p);
    else
      layernorm_bwd_kernel<NITER, false><<<grid, kRowThreads, smem, stream>>>(*p);
  });
  GVK_CHECK_LAUNCH("layernorm_bwd");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) small_wgrad_kernel(const float* __restrict__ a, int lda, int ra, const float* __restrict__ b, int ldb, int rb, int M,
                                                            int rows_per_cta, float* __restrict__ dw) {
  __shared__ float sa[16][96], sb[16][96];
  const int tid = threadIdx.x;
  const int nout = ra * rb;
  float acc[16];  // outputs tid, tid+256, ... (ra*rb <= 4096)
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const int m_begin = blockIdx.x * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
  for (int m0 = m_begin; m0 < m_end; m0 += 16) {
    __syncthreads();
    for (int idx = tid; idx < 16 * 96; idx += 256) {
      const int rr = idx / 96, j = idx - rr * 96;
      const int m = m0 + rr;
      sa[rr][j] = (m < m_end && j < ra) ? a[(size_t)m * lda + j] : 0.f;
      sb[rr][j] = (m < m_end && j < rb) ? b[(size_t)m * ldb + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int o = tid + 256 * i;
      if (o < nout) {
        const int j = o / rb, k = o - j * rb;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) acc[i] = fmaf(sa[rr][j], sb[rr][k], acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int o = tid + 256 * i;
    if (o < nout) atomicAdd(dw + o, acc[i]);
  }
}

int small_wgrad(const float* a, int lda, int ra, const float* b, int ldb, int rb, int M, float* dw, cudaStream_t stream) {
  GVK_CHECK_ARG(a && b && dw && M > 0, "gvk_small_wgrad: null pointer");
  GVK_CHECK_ARG(ra >= 1 && ra <= 96 && rb >= 1 && rb <= 96 && ra * rb <= 4096, "gvk_small_wgrad: ra=%d rb=%d must be in [1,96], ra*rb <= 4096", ra, rb);
  const int ctas = std::max(1, std::min(sm_count() * 2, (M + 127) / 128));
  int rows_per_cta = ((M + ctas - 1) / ctas + 15) / 16 * 16;
  const int grid = (M + rows_per_cta - 1) / rows_per_cta;
  small_wgrad_kernel<<<grid, 256, 0, stream>>>(a, lda, ra, b, ldb, rb, M, rows_per_cta, dw);
  GVK_CHECK_LAUNCH("small_wgrad");
  return GVK_OK;
}

// out[m, k] = sum_j a[m, j] * w[j * rb + k]
__global__ void __launch_bounds__(256) small_matmul_kernel(const float* __restrict__ a, int lda, int ra, const float* __restrict__ w, int rb, int M,
                                                             float* __restrict__ out, int ldo) {
  __shared__ float sw[64 * 64];
  for (int i = threadIdx.x; i < ra * rb; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const size_t total = (size_t)M * rb;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / rb), k = (int)(i - (size_t)m * rb);
    const float* ar = a + (size_t)m * lda;
    float s = 0.f;
    for (int j = 0; j < ra; ++j) s = fmaf(ar[j], sw[j * rb + k], s);
    out[(size_t)m * ldo + k] = s;
  }
}

int small_matmul(const float* a, int lda, int ra, const float* w, int rb, int M, float* out, int ldo, cudaStream_t stream) {
  GVK_CHECK_ARG(a && w && out && M > 0, "gvk_small_matmul: null pointer");
  GVK_CHECK_ARG(ra >= 1 && ra <= 96 && rb >= 1 && rb <= 96 && ra * rb <= 4096, "gvk_small_matmul: ra=%d rb=%d must be in [1,96], ra*rb <= 4096", ra, rb);
  const size_t total = (size_t)M * rb;
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 8);
  small_matmul_kernel<<<grid, 256, 0, stream>>>(a, lda, ra, w, rb, M, out, ldo);
  GVK_CHECK_LAUNCH("small_matmul");
  return GVK_OK;
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int ldx, int M, int dim, int rows_per_cta, float* __restrict__ out) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= dim) return;
  const int m_begin = blockIdx.x * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
  float s = 0.f;
  for (int m = m_begin; m < m_end; ++m) s += x[(size_t)m * ldx + c];
  atomicAdd(out + c, s);
}

int colsum(const float* x, int ldx, int M, int dim, float* out, cudaStream_t stream) {
  GVK_CHECK_ARG(x && out && M > 0 && dim > 0, "gvk_colsum: bad argument");
  const int ctas = std::max(1, std::min(sm_count() * 4, (M + 31) / 32));
  const int rows_per_cta = (M + ctas - 1) / ctas;
  dim3 grid((M + rows_per_cta - 1) / rows_per_cta, (dim + 255) / 256);
  colsum_kernel<<<grid, 256, 0, stream>>>(x, ldx, M, dim, rows_per_cta, out);
  GVK_CHECK_LAUNCH("colsum");
  return GVK_OK;
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy, int M, int dim2) {
  const size_t total = (size_t)M * dim2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / dim2), c = (int)(i - (size_t)m * dim2) * 2;
    const float2 v = *reinterpret_cast<const float2*>(x + (size_t)m * ldx + c);
    *reinterpret_cast<__nv_bfloat162*>(y + (size_t)m * ldy + c) = __floats2bfloat162_rn(v.x, v.y);
  }
}

int cast_f32_bf16(const float* x, int ldx, void* y, int ldy, int M, int dim, cudaStream_t stream) {
  GVK_CHECK_ARG(x && y && M > 0 && dim > 0 && dim % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0, "gvk_cast_f32_bf16: bad argument");
  const size_t total = (size_t)M * (dim / 2);
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
  cast_f32_bf16_kernel<<<grid, 256, 0, stream>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, M, dim / 2);
  GVK_CHECK_LAUNCH("cast_f32_bf16");
  return GVK_OK;
}

}  // namespace gvk
