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
// Register-blocked rank-r kernels.  A warp owns ROWS token rows at a time (ROWS = 4, or 2 for dim = 1024) so every weight value
// fetched from shared memory feeds ROWS FMAs (the one-row-per-warp form is bound by the LDS pipe, not by HBM), and all of the
// group's global loads are issued before the first use (ROWS x 3 KB in flight per warp).
// ------------------------------------------------------------------------------------------------
template <int N>
struct ReduceStep {
  // Halving exchange: lanes with bit `o` set keep the upper half of v[0, 2N), the others the lower half; the partner's copy is added.
  static __device__ __forceinline__ void run(float* v, int lane, int o) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float keep = upper ? v[k + N] : v[k];
      const float send = upper ? v[k] : v[k + N];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
};
// Sum 96 per-lane partials across the warp with 93 shuffles (a plain butterfly needs 480).  On return lane L holds in v[0..2] the
// totals of indices base, base+1, base+2 with base = 48 b4 + 24 b3 + 12 b2 + 6 b1 + 3 b0 (b_k = bit k of L).
__device__ __forceinline__ void warp_reduce_scatter96(float (&v)[96], int lane) {
  ReduceStep<48>::run(v, lane, 16);
  ReduceStep<24>::run(v, lane, 8);
  ReduceStep<12>::run(v, lane, 4);
  ReduceStep<6>::run(v, lane, 2);
  ReduceStep<3>::run(v, lane, 1);
}
__device__ __forceinline__ int reduce_scatter96_base(int lane) {
  return 48 * ((lane >> 4) & 1) + 24 * ((lane >> 3) & 1) + 12 * ((lane >> 2) & 1) + 6 * ((lane >> 1) & 1) + 3 * (lane & 1);
}

template <int NITER>
struct RowBlock {
  static constexpr int ROWS = NITER <= 12 ? 4 : 2;
  static constexpr int JCH = 96 / ROWS;  // projection outputs handled per pass over the row registers
};

static inline int row_block_grid(int M, int rows_per_warp) {
  const int blocks = (M + kRowWarps * rows_per_warp - 1) / (kRowWarps * rows_per_warp);
  return std::max(1, std::min(blocks, sm_count()));
}

// ------------------------------------------------------------------------------------------------
// rank-r down projection (+ optional dropout / LN prologue, activation, chained second projection)
// ------------------------------------------------------------------------------------------------
template <int NITER>
__global__ void __launch_bounds__(kRowThreads, 1) rowproj_down_kernel(gvk_rowproj_down_params p, int rpad) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  extern __shared__ float smem[];
  constexpr int dim = NITER * 64;
  constexpr int ROWS = RowBlock<NITER>::ROWS, JCH = RowBlock<NITER>::JCH;
  float* sw = smem;                       // [rpad][dim], rows >= r are zero
  float* sw2 = sw + (size_t)rpad * dim;   // [r2][r]
  float* sz = sw2 + p.r2 * p.r;           // [kRowWarps][ROWS][rpad] activated latents (chained projection only)
  stage_weight(sw, p.w, p.r, dim, p.w_sj, p.w_sc);
  for (int idx = p.r * dim + threadIdx.x; idx < rpad * dim; idx += blockDim.x) sw[idx] = 0.f;
  if (p.w2)
    for (int idx = threadIdx.x; idx < p.r2 * p.r; idx += blockDim.x) sw2[idx] = p.w2[idx];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_dim = 1.0f / dim;
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  const int base = reduce_scatter96_base(lane);
  const int q_out = base / JCH, j_out = base % JCH;   // this lane's outputs after the reduce: row q_out, latents j_out .. j_out + 2
  float* szw = sz + (size_t)warp * ROWS * rpad;
  const int ngroups = (p.M + ROWS - 1) / ROWS;
  for (int grp = blockIdx.x * kRowWarps + warp; grp < ngroups; grp += gridDim.x * kRowWarps) {
    const int row0 = grp * ROWS;
    float2 xv[ROWS][NITER];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      const float* xr = p.x + (size_t)min(row0 + q, p.M - 1) * p.ldx + lane * 2;
#pragma unroll
      for (int i = 0; i < NITER; ++i) xv[q][i] = *reinterpret_cast<const float2*>(xr + 64 * i);
    }
    if (p.drop_p > 0.f) {
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        const uint64_t e0 = p.offset + (uint64_t)min(row0 + q, p.M - 1) * dim + lane * 2;
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          const float2 m = drop_mult2(seed_eff, e0 + 64 * i, p.drop_p, inv_keep);
          xv[q][i].x *= m.x;
          xv[q][i].y *= m.y;
        }
      }
    }
    if (p.ln_gamma) {
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NITER; ++i) s += xv[q][i].x + xv[q][i].y;
        const float mean = warp_sum(s) * inv_dim;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          const float a = xv[q][i].x - mean, b = xv[q][i].y - mean;
          v += a * a + b * b;
        }
        const float rstd = rsqrtf(warp_sum(v) * inv_dim + p.eps);
        if (lane == 0 && row0 + q < p.M) {
          if (p.mean) p.mean[row0 + q] = mean;
          if (p.rstd) p.rstd[row0 + q] = rstd;
        }
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          const int c = lane * 2 + 64 * i;
          const float2 g = *reinterpret_cast<const float2*>(p.ln_gamma + c);
          const float2 b = *reinterpret_cast<const float2*>(p.ln_beta + c);
          xv[q][i].x = (xv[q][i].x - mean) * rstd * g.x + b.x;
          xv[q][i].y = (xv[q][i].y - mean) * rstd * g.y + b.y;
        }
      }
    }
    const int orow = row0 + q_out;
    for (int j0 = 0; j0 < p.r; j0 += JCH) {
      float acc[96];
#pragma unroll
      for (int jj = 0; jj < JCH; ++jj) {
        const float* wj = sw + (size_t)(j0 + jj) * dim + lane * 2;
        float a[ROWS];
#pragma unroll
        for (int q = 0; q < ROWS; ++q) a[q] = 0.f;
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          const float2 w = *reinterpret_cast<const float2*>(wj + 64 * i);
#pragma unroll
          for (int q = 0; q < ROWS; ++q) a[q] = fmaf(xv[q][i].y, w.y, fmaf(xv[q][i].x, w.x, a[q]));
        }
#pragma unroll
        for (int q = 0; q < ROWS; ++q) acc[q * JCH + jj] = a[q];
      }
      warp_reduce_scatter96(acc, lane);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int j = j0 + j_out + u;
        if (j < p.r) {
          float pre = acc[u] + (p.bias ? p.bias[j] : 0.f);
          if (orow < p.M && p.pre) p.pre[(size_t)orow * p.ldz + j] = pre;
          if (p.act == GVK_ROWACT_QUICKGELU)
            pre = quick_gelu(pre);
          else if (p.act == GVK_ROWACT_RELU)
            pre = fmaxf(pre, 0.f);
          if (orow < p.M) p.z[(size_t)orow * p.ldz + j] = pre;
          if (p.w2) szw[q_out * rpad + j] = pre;
        }
      }
    }
    if (p.w2) {
      __syncwarp();
      for (int o = lane; o < ROWS * p.r2; o += 32) {
        const int q = o / p.r2, k = o - q * p.r2;
        const float* zq = szw + q * rpad;
        const float* wk = sw2 + k * p.r;
        float s = 0.f;
        for (int j = 0; j < p.r; ++j) s = fmaf(zq[j], wk[j], s);
        if (row0 + q < p.M) p.z2[(size_t)(row0 + q) * p.ldz2 + k] = s;
      }
      __syncwarp();
    }
  }
}

static size_t rowproj_down_smem(int r, int r2, int dim, int* rpad_out) {
  const int jch = dim <= 768 ? 24 : 48, rows = dim <= 768 ? 4 : 2;
  const int rpad = (r + jch - 1) / jch * jch;
  *rpad_out = rpad;
  return ((size_t)rpad * dim + (size_t)r2 * r + (r2 > 0 ? (size_t)kRowWarps * rows * rpad : 0)) * sizeof(float);
}

int rowproj_down(const gvk_rowproj_down_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->x && p->w && p->z, "gvk_rowproj_down: null pointer");
  GVK_CHECK_ARG(p->r >= 1 && p->r <= 96 && p->r2 >= 0 && p->r2 <= 96, "gvk_rowproj_down: r=%d (<=96), r2=%d (<=96)", p->r, p->r2);
  GVK_CHECK_ARG(p->M > 0 && p->ldx % 2 == 0, "gvk_rowproj_down: bad shape");
  GVK_CHECK_ARG(!p->w2 || p->z2, "gvk_rowproj_down: chained projection needs z2");
  if (p->precision == GVK_PREC_TF32) return rowproj_down_tc(p, stream);
  int rpad = 0;
  const size_t smem = rowproj_down_smem(p->r, p->w2 ? p->r2 : 0, p->dim, &rpad);
  if (smem > 227 * 1024) {
    set_last_error("gvk_rowproj_down: r=%d at dim=%d needs %zu B of shared memory (> 227 KB)", p->r, p->dim, smem);
    return GVK_ERR_UNSUPPORTED;
  }
  GVK_DISPATCH_NITER(p->dim, {
    static size_t configured = 0;
    if (smem > configured) {
      int st = cuda_status(cudaFuncSetAttribute(rowproj_down_kernel<NITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "rowproj_down smem");
      if (st != GVK_OK) return st;
      configured = smem;
    }
    rowproj_down_kernel<NITER><<<row_block_grid(p->M, RowBlock<NITER>::ROWS), kRowThreads, smem, stream>>>(*p, rpad);
  });
  GVK_CHECK_LAUNCH("rowproj_down");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// rank-r up projection + bias + dropout + residual
// ------------------------------------------------------------------------------------------------
template <int NITER>
__global__ void __launch_bounds__(kRowThreads, 1) rowproj_up_kernel(gvk_rowproj_up_params p) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  extern __shared__ float smem[];
  constexpr int dim = NITER * 64;
  constexpr int ROWS = RowBlock<NITER>::ROWS;
  float* sw = smem;  // [r][dim]
  stage_weight(sw, p.w, p.r, dim, p.w_sj, p.w_sc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  const bool early_res = p.res && !(p.drop_p > 0.f);
  const int ngroups = (p.M + ROWS - 1) / ROWS;
  for (int grp = blockIdx.x * kRowWarps + warp; grp < ngroups; grp += gridDim.x * kRowWarps) {
    const int row0 = grp * ROWS;
    float cl[ROWS][3];   // lane holds latents lane, lane + 32, lane + 64 of each row
    float2 acc[ROWS][NITER];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      const size_t row = (size_t)min(row0 + q, p.M - 1);
#pragma unroll
      for (int u = 0; u < 3; ++u) cl[q][u] = (lane + 32 * u < p.r) ? p.c[row * p.ldc + lane + 32 * u] : 0.f;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const int c = lane * 2 + 64 * i;
        float2 a = p.bias ? *reinterpret_cast<const float2*>(p.bias + c) : make_float2(0.f, 0.f);
        if (early_res) {
          const float2 r = *reinterpret_cast<const float2*>(p.res + row * p.ld_res + c);
          a.x += r.x;
          a.y += r.y;
        }
        acc[q][i] = a;
      }
    }
    for (int j = 0; j < p.r; ++j) {
      float cj[ROWS];
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        const float src = j < 32 ? cl[q][0] : (j < 64 ? cl[q][1] : cl[q][2]);
        cj[q] = __shfl_sync(0xffffffffu, src, j & 31);
      }
      const float* wj = sw + (size_t)j * dim + lane * 2;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const float2 w = *reinterpret_cast<const float2*>(wj + 64 * i);
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
          acc[q][i].x = fmaf(cj[q], w.x, acc[q][i].x);
          acc[q][i].y = fmaf(cj[q], w.y, acc[q][i].y);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      const int row = row0 + q;
      if (row >= p.M) break;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const int c = lane * 2 + 64 * i;
        float2 v = acc[q][i];
        if (p.drop_p > 0.f) {
          const float2 m = drop_mult2(seed_eff, p.offset + (uint64_t)row * dim + c, p.drop_p, inv_keep);
          v.x *= m.x;
          v.y *= m.y;
          if (p.res) {
            const float2 r = *reinterpret_cast<const float2*>(p.res + (size_t)row * p.ld_res + c);
            v.x += r.x;
            v.y += r.y;
          }
        }
        *reinterpret_cast<float2*>(p.out + (size_t)row * p.ld_out + c) = v;
        if (p.out_lp) *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.out_lp) + (size_t)row * p.ld_out_lp + c) = __floats2bfloat162_rn(v.x, v.y);
      }
    }
  }
}

int rowproj_up(const gvk_rowproj_up_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->c && p->w && p->out, "gvk_rowproj_up: null pointer");
  GVK_CHECK_ARG(p->r >= 1 && p->r <= 96, "gvk_rowproj_up: r=%d must be in [1,96]", p->r);
  GVK_CHECK_ARG(p->M > 0 && p->ld_out % 2 == 0 && (!p->res || p->ld_res % 2 == 0), "gvk_rowproj_up: bad shape");
  if (p->precision == GVK_PREC_TF32) return rowproj_up_tc(p, stream);
  const size_t smem = (size_t)p->r * p->dim * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("gvk_rowproj_up: r=%d at dim=%d needs %zu B of shared memory (> 227 KB)", p->r, p->dim, smem);
    return GVK_ERR_UNSUPPORTED;
  }
  GVK_DISPATCH_NITER(p->dim, {
    static size_t configured = 0;
    if (smem > configured) {
      int st = cuda_status(cudaFuncSetAttribute(rowproj_up_kernel<NITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "rowproj_up smem");
      if (st != GVK_OK) return st;
      configured = smem;
    }
    rowproj_up_kernel<NITER><<<row_block_grid(p->M, RowBlock<NITER>::ROWS), kRowThreads, smem, stream>>>(*p);
  });
  GVK_CHECK_LAUNCH("rowproj_up");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// rank-r weight gradient  dw(j,c) += sum_m a[m,j] * f(x[m,c])
// A thread owns NCOL columns and all r latents of them (r*NCOL accumulators); rows are consumed kWgRows at a time with every
// global load of the chunk issued before the FMAs (clamped row index, zeroed latent for the tail: no break in the load path).
// ------------------------------------------------------------------------------------------------
constexpr int kWgRows = 8;  // rows staged per step

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) skinny_wgrad_kernel(gvk_skinny_wgrad_params p, int rows_per_cta) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  __shared__ __align__(16) float sa[kWgRows][R];
  __shared__ float smean[kWgRows], srstd[kWgRows];
  const int tid = threadIdx.x;
  const int m_begin = blockIdx.x * rows_per_cta;
  const int m_end = min(p.M, m_begin + rows_per_cta);
  const bool colv = tid * 4 < p.dim;            // thread owns columns 4 tid .. 4 tid + 3 (dim % 4 == 0)
  const int c0 = colv ? tid * 4 : 0;
  float acc[4][R];
  float xsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;
  float asum = 0.f;  // thread j < r
  float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.ln_gamma) {
    g = *reinterpret_cast<const float4*>(p.ln_gamma + c0);
    b = *reinterpret_cast<const float4*>(p.ln_beta + c0);
  }
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  for (int m0 = m_begin; m0 < m_end; m0 += kWgRows) {
    float4 xr[kWgRows];
#pragma unroll
    for (int rr = 0; rr < kWgRows; ++rr) xr[rr] = *reinterpret_cast<const float4*>(p.x + (size_t)min(m0 + rr, m_end - 1) * p.ldx + c0);
    __syncthreads();   // previous chunk's sa fully consumed
    for (int idx = tid; idx < kWgRows * R; idx += THREADS) {
      const int rr = idx / R, j = idx - rr * R;
      const int m = m0 + rr;
      sa[rr][j] = (m < m_end && j < p.r) ? p.a[(size_t)m * p.lda + j] : 0.f;   // zero latent: tail rows contribute nothing to dw
    }
    if (tid < kWgRows) {
      const int m = min(m0 + tid, m_end - 1);
      smean[tid] = p.mean ? p.mean[m] : 0.f;
      srstd[tid] = p.rstd ? p.rstd[m] : 1.f;
    }
    __syncthreads();
    if (tid < R) {
#pragma unroll
      for (int rr = 0; rr < kWgRows; ++rr) asum += sa[rr][tid];
    }
#pragma unroll
    for (int rr = 0; rr < kWgRows; ++rr) {
      float x[4] = {xr[rr].x, xr[rr].y, xr[rr].z, xr[rr].w};
      if (p.drop_p > 0.f) {
        const uint64_t e = p.offset + (uint64_t)min(m0 + rr, m_end - 1) * p.dim + c0;
        const float2 ma = drop_mult2(seed_eff, e, p.drop_p, inv_keep), mb = drop_mult2(seed_eff, e + 2, p.drop_p, inv_keep);
        x[0] *= ma.x; x[1] *= ma.y; x[2] *= mb.x; x[3] *= mb.y;
      }
      if (p.ln_gamma) {
        const float mu = smean[rr], rs = srstd[rr];
        x[0] = (x[0] - mu) * rs * g.x + b.x;
        x[1] = (x[1] - mu) * rs * g.y + b.y;
        x[2] = (x[2] - mu) * rs * g.z + b.z;
        x[3] = (x[3] - mu) * rs * g.w + b.w;
      }
      if (m0 + rr < m_end) {
#pragma unroll
        for (int i = 0; i < 4; ++i) xsum[i] += x[i];
      }
#pragma unroll
      for (int j = 0; j < R; j += 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sa[rr][j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][j] = fmaf(a4.x, x[i], acc[i][j]);
          acc[i][j + 1] = fmaf(a4.y, x[i], acc[i][j + 1]);
          acc[i][j + 2] = fmaf(a4.z, x[i], acc[i][j + 2]);
          acc[i][j + 3] = fmaf(a4.w, x[i], acc[i][j + 3]);
        }
      }
    }
  }
  // per-CTA partials (no atomics: same-address contention from ~300 CTAs serialises in L2); reduced by skinny_wgrad_reduce_kernel
  if (colv) {
    float* ws_dw = p.ws + (size_t)blockIdx.x * p.r * p.dim;
    float* ws_dx = p.ws + (size_t)gridDim.x * p.r * p.dim + (size_t)blockIdx.x * p.dim;
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (j < p.r) *reinterpret_cast<float4*>(ws_dw + (size_t)j * p.dim + c0) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    *reinterpret_cast<float4*>(ws_dx + c0) = make_float4(xsum[0], xsum[1], xsum[2], xsum[3]);
  }
  if (tid < 32) {
    float* ws_da = p.ws + (size_t)gridDim.x * (p.r + 1) * p.dim + (size_t)blockIdx.x * 32;
    ws_da[tid] = tid < p.r ? asum : 0.f;
  }
}

// Sum the per-CTA partials: a block owns 32 consecutive outputs, its 16 warps each sum every 16th CTA's partial (coalesced 128-byte
// reads, independent loads in flight), then a fixed-order shared-memory tree adds the 16 subtotals: deterministic, ~500 blocks.
constexpr int kRedSplit = 16;
__global__ void __launch_bounds__(32 * kRedSplit) skinny_wgrad_reduce_kernel(gvk_skinny_wgrad_params p, int ncta) {
  __shared__ float part[kRedSplit][33];
  const int rd = p.r * p.dim;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;
  const int total = rd + p.dim + 32;
  float s = 0.f;
  if (idx < total) {
    const float* base;
    size_t stride;
    if (idx < rd) { base = p.ws + idx; stride = rd; }
    else if (idx < rd + p.dim) { base = p.ws + (size_t)ncta * rd + (idx - rd); stride = p.dim; }
    else { base = p.ws + (size_t)ncta * (p.r + 1) * p.dim + (idx - rd - p.dim); stride = 32; }
    float s0 = 0.f, s1 = 0.f;
    int k = grp;
    for (; k + kRedSplit < ncta; k += 2 * kRedSplit) {
      s0 += base[(size_t)k * stride];
      s1 += base[(size_t)(k + kRedSplit) * stride];
    }
    if (k < ncta) s0 += base[(size_t)k * stride];
    s = s0 + s1;
  }
  part[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && idx < total) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kRedSplit; ++w) tot += part[w][lane];
    if (idx < rd) {
      if (p.dw) {
        const int j = idx / p.dim, c = idx - j * p.dim;
        p.dw[(size_t)j * p.dw_sj + (size_t)c * p.dw_sc] += tot;
      }
    } else if (idx < rd + p.dim) {
      if (p.dx_colsum) p.dx_colsum[idx - rd] += tot;
    } else {
      const int j = idx - rd - p.dim;
      if (p.da_colsum && j < p.r) p.da_colsum[j] += tot;
    }
  }
}

void skinny_wgrad_plan(int M, int* ctas, int* rows_per_cta) {
  const int want = std::max(1, std::min(sm_count() * 2, (M + 31) / 32));
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

void skinny_wgrad_launch_reduce(const gvk_skinny_wgrad_params* p, int ncta, cudaStream_t stream) {
  const int total = p->r * p->dim + p->dim + 32;
  skinny_wgrad_reduce_kernel<<<(total + 31) / 32, 32 * kRedSplit, 0, stream>>>(*p, ncta);
}

template <int THREADS>
static int skinny_wgrad_launch(const gvk_skinny_wgrad_params* p, cudaStream_t stream) {
  int grid, rows_per_cta;
  skinny_wgrad_plan(p->M, &grid, &rows_per_cta);
  if (p->r <= 8)
    skinny_wgrad_kernel<8, THREADS><<<grid, THREADS, 0, stream>>>(*p, rows_per_cta);
  else if (p->r <= 20)
    skinny_wgrad_kernel<20, THREADS><<<grid, THREADS, 0, stream>>>(*p, rows_per_cta);
  else
    skinny_wgrad_kernel<32, THREADS><<<grid, THREADS, 0, stream>>>(*p, rows_per_cta);
  GVK_CHECK_LAUNCH("skinny_wgrad");
  skinny_wgrad_launch_reduce(p, grid, stream);
  GVK_CHECK_LAUNCH("skinny_wgrad_reduce");
  return GVK_OK;
}

int skinny_wgrad(const gvk_skinny_wgrad_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->a && p->x, "gvk_skinny_wgrad: null pointer");
  GVK_CHECK_ARG(p->r >= 1 && p->r <= 32 && p->dim >= 4 && p->dim <= 1024 && p->dim % 4 == 0 && p->ldx % 4 == 0 && p->M > 0,
                "gvk_skinny_wgrad: r=%d dim=%d ldx=%d M=%d (r <= 32, dim and ldx multiples of 4, dim <= 1024)", p->r, p->dim, p->ldx, p->M);
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(p->x) & 15) == 0, "gvk_skinny_wgrad: x must be 16-byte aligned");
  GVK_CHECK_ARG(!p->ln_gamma || (p->ln_beta && p->mean && p->rstd), "gvk_skinny_wgrad: LN recompute needs beta, mean, rstd");
  GVK_CHECK_ARG(p->ws && p->ws_floats >= skinny_wgrad_ws_floats(p->r, p->dim, p->M), "gvk_skinny_wgrad: workspace of %zu floats required (gvk_skinny_wgrad_ws_floats)",
                skinny_wgrad_ws_floats(p->r, p->dim, p->M));
  if (p->precision == GVK_PREC_TF32 && skinny_wgrad_tc_supported(p)) return skinny_wgrad_tc(p, stream);
  // thread = 4 adjacent columns x all r latents (one 16-byte load per row); 2+ CTAs per SM keep >= 48 KB of loads in flight
  if (p->dim <= 256) return skinny_wgrad_launch<64>(p, stream);
  if (p->dim <= 384) return skinny_wgrad_launch<96>(p, stream);
  if (p->dim <= 768) return skinny_wgrad_launch<192>(p, stream);
  return skinny_wgrad_launch<256>(p, stream);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward
//   dx = dres + LN'(dy) + az @ aw        dy dense, or rank-r (dy = dz @ w);  az @ aw is an optional rank-ra term added OUTSIDE the
//   norm (the dgrad of a down-projection that reads the same residual stream: model/gaviko.py:155 after :304's LayerNorm input).
// Two rows per warp so the rank-r panels read from shared memory feed two rows of FMAs, and both rows' loads are in flight together.
// ------------------------------------------------------------------------------------------------
template <int NITER, bool PGRAD>
__global__ void __launch_bounds__(kRowThreads, 1) layernorm_bwd_kernel(gvk_layernorm_bwd_params p) {
  extern __shared__ float smem[];
  constexpr int dim = NITER * 64;
  constexpr int ROWS = 2;
  float* sw = smem;                                   // [r][dim] when dy is given in rank-r form
  float* saw = sw + (p.dz ? (size_t)p.r * dim : 0);   // [ra][dim] additive rank-ra term
  float* red = saw + (p.az ? (size_t)p.ra * dim : 0); // [kRowWarps][2*dim] parameter-gradient staging
  if (p.dz) stage_weight(sw, p.w, p.r, dim, p.w_sj, p.w_sc);
  if (p.az) stage_weight(saw, p.aw, p.ra, dim, p.aw_sj, p.aw_sc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_dim = 1.0f / dim;
  constexpr int NG = PGRAD ? NITER : 1;
  float2 dg[NG], db[NG], dsc[NG], dsh[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) dg[i] = db[i] = dsc[i] = dsh[i] = make_float2(0.f, 0.f);
  const int ngroups = (p.M + ROWS - 1) / ROWS;
  for (int grp = blockIdx.x * kRowWarps + warp; grp < ngroups; grp += gridDim.x * kRowWarps) {
    const int row0 = grp * ROWS;
    size_t rowc[ROWS];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) rowc[q] = (size_t)min(row0 + q, p.M - 1);
    float2 dy[ROWS][NITER], xh[ROWS][NITER];
    // ---- issue the streaming loads first
#pragma unroll
    for (int q = 0; q < ROWS; ++q)
#pragma unroll
      for (int i = 0; i < NITER; ++i) xh[q][i] = *reinterpret_cast<const float2*>(p.x + rowc[q] * p.ldx + lane * 2 + 64 * i);
    if (p.dz) {
      float zl[ROWS];
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        zl[q] = lane < p.r ? p.dz[rowc[q] * p.ld_dz + lane] : 0.f;
#pragma unroll
        for (int i = 0; i < NITER; ++i) dy[q][i] = make_float2(0.f, 0.f);
      }
      for (int j = 0; j < p.r; ++j) {
        float zj[ROWS];
#pragma unroll
        for (int q = 0; q < ROWS; ++q) zj[q] = __shfl_sync(0xffffffffu, zl[q], j);
        const float* wj = sw + (size_t)j * dim + lane * 2;
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          const float2 w = *reinterpret_cast<const float2*>(wj + 64 * i);
#pragma unroll
          for (int q = 0; q < ROWS; ++q) {
            dy[q][i].x = fmaf(zj[q], w.x, dy[q][i].x);
            dy[q][i].y = fmaf(zj[q], w.y, dy[q][i].y);
          }
        }
      }
    }
    if (p.dy) {
#pragma unroll
      for (int q = 0; q < ROWS; ++q)
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          float2 d;
          if (p.dy_dtype == GVK_BF16)
            d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + rowc[q] * p.ld_dy + lane * 2 + 64 * i));
          else
            d = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p.dy) + rowc[q] * p.ld_dy + lane * 2 + 64 * i);
          if (p.dz) {
            dy[q][i].x += d.x;
            dy[q][i].y += d.y;
          } else {
            dy[q][i] = d;
          }
        }
    }
    float m1[ROWS], m2[ROWS], rstd[ROWS];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      const float mean = p.mean[rowc[q]];
      rstd[q] = p.rstd[rowc[q]];
      const bool live = row0 + q < p.M;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const float2 gam = *reinterpret_cast<const float2*>(p.gamma + lane * 2 + 64 * i);
        xh[q][i] = make_float2((xh[q][i].x - mean) * rstd[q], (xh[q][i].y - mean) * rstd[q]);
        if constexpr (PGRAD) {
          if (p.ssf_scale) {   // y = (xh * gamma + beta) * ssf_scale + ssf_shift: reduce the SSF gradients, then continue with dy * ssf_scale
            const int c = lane * 2 + 64 * i;
            const float2 bt = *reinterpret_cast<const float2*>(p.beta + c);
            const float2 sc = *reinterpret_cast<const float2*>(p.ssf_scale + c);
            if (live) {
              dsc[i].x += dy[q][i].x * (xh[q][i].x * gam.x + bt.x);
              dsc[i].y += dy[q][i].y * (xh[q][i].y * gam.y + bt.y);
              dsh[i].x += dy[q][i].x;
              dsh[i].y += dy[q][i].y;
            }
            dy[q][i].x *= sc.x;
            dy[q][i].y *= sc.y;
          }
          if (live) {
            dg[i].x += dy[q][i].x * xh[q][i].x;
            dg[i].y += dy[q][i].y * xh[q][i].y;
            db[i].x += dy[q][i].x;
            db[i].y += dy[q][i].y;
          }
        }
        dy[q][i].x *= gam.x;  // g = dy * gamma
        dy[q][i].y *= gam.y;
        s1 += dy[q][i].x + dy[q][i].y;
        s2 += dy[q][i].x * xh[q][i].x + dy[q][i].y * xh[q][i].y;
      }
      m1[q] = warp_sum(s1) * inv_dim;
      m2[q] = warp_sum(s2) * inv_dim;
    }
    // dx (before the additive terms) overwrites dy
#pragma unroll
    for (int q = 0; q < ROWS; ++q)
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        dy[q][i].x = rstd[q] * (dy[q][i].x - m1[q] - xh[q][i].x * m2[q]);
        dy[q][i].y = rstd[q] * (dy[q][i].y - m1[q] - xh[q][i].y * m2[q]);
      }
    if (p.dres) {
#pragma unroll
      for (int q = 0; q < ROWS; ++q)
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          const float2 r = *reinterpret_cast<const float2*>(p.dres + rowc[q] * p.ld_dres + lane * 2 + 64 * i);
          dy[q][i].x += r.x;
          dy[q][i].y += r.y;
        }
    }
    if (p.az) {
      float zl[ROWS];
#pragma unroll
      for (int q = 0; q < ROWS; ++q) zl[q] = lane < p.ra ? p.az[rowc[q] * p.ld_az + lane] : 0.f;
      for (int j = 0; j < p.ra; ++j) {
        float zj[ROWS];
#pragma unroll
        for (int q = 0; q < ROWS; ++q) zj[q] = __shfl_sync(0xffffffffu, zl[q], j);
        const float* wj = saw + (size_t)j * dim + lane * 2;
#pragma unroll
        for (int i = 0; i < NITER; ++i) {
          const float2 w = *reinterpret_cast<const float2*>(wj + 64 * i);
#pragma unroll
          for (int q = 0; q < ROWS; ++q) {
            dy[q][i].x = fmaf(zj[q], w.x, dy[q][i].x);
            dy[q][i].y = fmaf(zj[q], w.y, dy[q][i].y);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      const int row = row0 + q;
      if (row >= p.M) break;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const int c = lane * 2 + 64 * i;
        *reinterpret_cast<float2*>(p.dx + (size_t)row * p.ld_dx + c) = dy[q][i];
        if (p.dx_lp) *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.dx_lp) + (size_t)row * p.ld_dx_lp + c) = __floats2bfloat162_rn(dy[q][i].x, dy[q][i].y);
      }
    }
  }
  if constexpr (PGRAD) {
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
    if (p.ssf_scale) {   // second round through the same staging area for the SSF gradients
      __syncthreads();
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const int c = lane * 2 + 64 * i;
        red[warp * 2 * dim + c] = dsc[i].x;
        red[warp * 2 * dim + c + 1] = dsc[i].y;
        red[warp * 2 * dim + dim + c] = dsh[i].x;
        red[warp * 2 * dim + dim + c + 1] = dsh[i].y;
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < 2 * dim; idx += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kRowWarps; ++w) s += red[w * 2 * dim + idx];
        if (idx < dim) {
          if (p.dssf_scale) atomicAdd(p.dssf_scale + idx, s);
        } else {
          if (p.dssf_shift) atomicAdd(p.dssf_shift + idx - dim, s);
        }
      }
    }
  }
}

int layernorm_bwd(const gvk_layernorm_bwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->x && p->gamma && p->mean && p->rstd && p->dx, "gvk_layernorm_bwd: null pointer");
  GVK_CHECK_ARG(p->dy != nullptr || p->dz != nullptr, "gvk_layernorm_bwd: dy and / or dz must be given");
  GVK_CHECK_ARG(p->dy_dtype == GVK_F32 || p->dy_dtype == GVK_BF16, "gvk_layernorm_bwd: dy_dtype must be GVK_F32 or GVK_BF16");
  GVK_CHECK_ARG(p->dy_dtype == GVK_F32 || static_cast<const void*>(p->dx) != p->dy, "gvk_layernorm_bwd: dx cannot alias a bf16 dy");
  GVK_CHECK_ARG(!p->ssf_scale || p->beta, "gvk_layernorm_bwd: the SSF form needs beta (to rebuild the LayerNorm output)");
  GVK_CHECK_ARG(!p->dz || (p->w && p->r >= 1 && p->r <= 32), "gvk_layernorm_bwd: rank-r form needs w and 1 <= r <= 32");
  GVK_CHECK_ARG(!p->az || (p->aw && p->ra >= 1 && p->ra <= 32), "gvk_layernorm_bwd: additive rank term needs aw and 1 <= ra <= 32");
  GVK_CHECK_ARG(!p->ow || (p->oz && p->orank >= 1 && p->orank <= 32), "gvk_layernorm_bwd: output projection needs oz and 1 <= orank <= 32");
  if (p->precision == GVK_PREC_TF32 && layernorm_bwd_tc_supported(p)) return layernorm_bwd_tc(p, stream);
  if (p->ow) {
    set_last_error("gvk_layernorm_bwd: the output projection exists in the tensor-core form only (precision TF32, dense bf16 dy, no az / dz / dgamma / dbeta, dim 384 or 768)");
    return GVK_ERR_UNSUPPORTED;
  }
  const bool red = p->dgamma || p->dbeta || p->ssf_scale;
  const size_t smem = ((p->dz ? (size_t)p->r * p->dim : 0) + (p->az ? (size_t)p->ra * p->dim : 0) + (red ? (size_t)kRowWarps * 2 * p->dim : 0)) * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("gvk_layernorm_bwd: needs %zu B of shared memory (> 227 KB)", smem);
    return GVK_ERR_UNSUPPORTED;
  }
  const int grid = row_block_grid(p->M, 2);
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
// dw[j, k] += sum_m a[m, j] b[m, k]:  a thread owns up to 16 outputs (their operand columns are fixed for the whole kernel), the CTA
// stages 32 rows of both operands per step with 16-byte loads when the layout allows.
constexpr int kSwRows = 32;
__global__ void __launch_bounds__(256) small_wgrad_kernel(const float* __restrict__ a, int lda, int ra, const float* __restrict__ b, int ldb, int rb, int M,
                                                            int rows_per_cta, float* __restrict__ dw, int vec) {
  __shared__ __align__(16) float sa[kSwRows][96], sb[kSwRows][96];
  const int tid = threadIdx.x;
  const int nout = ra * rb;
  float acc[16];  // outputs tid, tid+256, ... (ra*rb <= 4096)
  int ja[16], kb[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    acc[i] = 0.f;
    const int o = min(tid + 256 * i, nout - 1);
    ja[i] = o / rb;
    kb[i] = o - ja[i] * rb;
  }
  const int nmine = (nout - tid + 255) / 256;   // outputs this thread owns (may be <= 0)
  const int m_begin = blockIdx.x * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
  for (int m0 = m_begin; m0 < m_end; m0 += kSwRows) {
    __syncthreads();
    if (vec) {
      const int va = ra / 4, vb = rb / 4;
      for (int idx = tid; idx < kSwRows * (va + vb); idx += 256) {
        const int rr = idx / (va + vb), c = idx - rr * (va + vb);
        const int m = m0 + rr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < va) {
          if (m < m_end) v = *reinterpret_cast<const float4*>(a + (size_t)m * lda + 4 * c);
          *reinterpret_cast<float4*>(&sa[rr][4 * c]) = v;
        } else {
          if (m < m_end) v = *reinterpret_cast<const float4*>(b + (size_t)m * ldb + 4 * (c - va));
          *reinterpret_cast<float4*>(&sb[rr][4 * (c - va)]) = v;
        }
      }
    } else {
      for (int idx = tid; idx < kSwRows * 96; idx += 256) {
        const int rr = idx / 96, j = idx - rr * 96;
        const int m = m0 + rr;
        sa[rr][j] = (m < m_end && j < ra) ? a[(size_t)m * lda + j] : 0.f;
        sb[rr][j] = (m < m_end && j < rb) ? b[(size_t)m * ldb + j] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < nmine) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int rr = 0; rr < kSwRows; rr += 2) {
          s0 = fmaf(sa[rr][ja[i]], sb[rr][kb[i]], s0);
          s1 = fmaf(sa[rr + 1][ja[i]], sb[rr + 1][kb[i]], s1);
        }
        acc[i] += s0 + s1;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < nmine) atomicAdd(dw + tid + 256 * i, acc[i]);
}

// Tensor-core form for ra <= 64, rb <= 24 (LocalSelfAttention.qkv: 60 x 20): D[j, k] += A^T[j, m] B[m, k] with the row index as the MMA k
// dimension, mma.sync m16n8k8 tf32 with both operands split into hi + lo (hi hi + lo hi + hi lo: ~2^-21 relative, inside the 1e-4 parity mode).
// The staged FMA form above reads two shared-memory words per FMA: 45 us for 64 k rows against 3 us of operand traffic.  Here a warp walks
// 8-row steps with its fragments loaded straight from global memory (the operands are a few MB and L2-resident), the 8 warps of a CTA meet in
// shared memory and the CTA issues one atomic per output.
constexpr int kSwtWarps = 8;
__global__ void __launch_bounds__(kSwtWarps * 32) small_wgrad_tc_kernel(const float* __restrict__ a, int lda, int ra, const float* __restrict__ b, int ldb, int rb, int M,
                                                                         float* __restrict__ dw) {
  constexpr int MT = 4, NT = 3;
  __shared__ float sacc[MT * 16 * NT * 8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < MT * 16 * NT * 8; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  const int nwarps = gridDim.x * kSwtWarps;
  for (int m0 = (blockIdx.x * kSwtWarps + warp) * 8; m0 < M; m0 += nwarps * 8) {
    const int r0 = m0 + t, r1 = m0 + t + 4;
    const bool v0 = r0 < M, v1 = r1 < M;
    float av[MT][4], bv[NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int j0 = 16 * mt + g, j1 = j0 + 8;
      av[mt][0] = (v0 && j0 < ra) ? a[(size_t)r0 * lda + j0] : 0.f;
      av[mt][1] = (v0 && j1 < ra) ? a[(size_t)r0 * lda + j1] : 0.f;
      av[mt][2] = (v1 && j0 < ra) ? a[(size_t)r1 * lda + j0] : 0.f;
      av[mt][3] = (v1 && j1 < ra) ? a[(size_t)r1 * lda + j1] : 0.f;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int k = 8 * nt + g;
      bv[nt][0] = (v0 && k < rb) ? b[(size_t)r0 * ldb + k] : 0.f;
      bv[nt][1] = (v1 && k < rb) ? b[(size_t)r1 * ldb + k] : 0.f;
    }
    uint32_t ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        ah[mt][e] = f2tf32(av[mt][e]);
        al[mt][e] = f2tf32(av[mt][e] - __uint_as_float(ah[mt][e]));
      }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        bh[nt][e] = f2tf32(bv[nt][e]);
        bl[nt][e] = f2tf32(bv[nt][e] - __uint_as_float(bh[nt][e]));
      }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        mma_tf32(acc[mt][nt], al[mt][0], al[mt][1], al[mt][2], al[mt][3], bh[nt][0], bh[nt][1]);
        mma_tf32(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bl[nt][0], bl[nt][1]);
        mma_tf32(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bh[nt][0], bh[nt][1]);
      }
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      atomicAdd(&sacc[(16 * mt + g) * (NT * 8) + 8 * nt + 2 * t], acc[mt][nt][0]);
      atomicAdd(&sacc[(16 * mt + g) * (NT * 8) + 8 * nt + 2 * t + 1], acc[mt][nt][1]);
      atomicAdd(&sacc[(16 * mt + g + 8) * (NT * 8) + 8 * nt + 2 * t], acc[mt][nt][2]);
      atomicAdd(&sacc[(16 * mt + g + 8) * (NT * 8) + 8 * nt + 2 * t + 1], acc[mt][nt][3]);
    }
  __syncthreads();
  for (int i = tid; i < ra * rb; i += blockDim.x) {
    const int j = i / rb, k = i - j * rb;
    atomicAdd(dw + i, sacc[j * (NT * 8) + k]);
  }
}

int small_wgrad(const float* a, int lda, int ra, const float* b, int ldb, int rb, int M, float* dw, cudaStream_t stream) {
  GVK_CHECK_ARG(a && b && dw && M > 0, "gvk_small_wgrad: null pointer");
  GVK_CHECK_ARG(ra >= 1 && ra <= 96 && rb >= 1 && rb <= 96 && ra * rb <= 4096, "gvk_small_wgrad: ra=%d rb=%d must be in [1,96], ra*rb <= 4096", ra, rb);
  if (ra <= 64 && rb <= 24) {
    const int grid = std::max(1, std::min(sm_count() * 2, (M + kSwtWarps * 8 - 1) / (kSwtWarps * 8)));
    small_wgrad_tc_kernel<<<grid, kSwtWarps * 32, 0, stream>>>(a, lda, ra, b, ldb, rb, M, dw);
    GVK_CHECK_LAUNCH("small_wgrad_tc");
    return GVK_OK;
  }
  const int ctas = std::max(1, std::min(sm_count() * 2, (M + 127) / 128));
  int rows_per_cta = ((M + ctas - 1) / ctas + kSwRows - 1) / kSwRows * kSwRows;
  const int grid = (M + rows_per_cta - 1) / rows_per_cta;
  const int vec = ra % 4 == 0 && rb % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  small_wgrad_kernel<<<grid, 256, 0, stream>>>(a, lda, ra, b, ldb, rb, M, rows_per_cta, dw, vec);
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
// Same product with a thread per row (all of its ra inputs in registers, 16-byte loads / stores) for the shapes of the window-attention
// backward (ra = 3 r, rb = r): the element-per-thread form above walks a dependent chain of ra scalar loads (34 us for 64 k rows).
template <int RA, int RB>
__global__ void __launch_bounds__(128) small_matmul_rows_kernel(const float* __restrict__ a, int lda, const float* __restrict__ w, int M, float* __restrict__ out, int ldo) {
  __shared__ __align__(16) float sw[RA * RB];
  for (int i = threadIdx.x; i < RA * RB; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float x[RA];
#pragma unroll
  for (int j = 0; j < RA; j += 4) {
    const float4 v = *reinterpret_cast<const float4*>(a + (size_t)m * lda + j);
    x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
  }
#pragma unroll
  for (int k = 0; k < RB; k += 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < RA; ++j) {
      const float4 wv = *reinterpret_cast<const float4*>(sw + j * RB + k);   // the same address in every lane: broadcast
      acc.x = fmaf(x[j], wv.x, acc.x); acc.y = fmaf(x[j], wv.y, acc.y); acc.z = fmaf(x[j], wv.z, acc.z); acc.w = fmaf(x[j], wv.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + (size_t)m * ldo + k) = acc;
  }
}

int small_matmul(const float* a, int lda, int ra, const float* w, int rb, int M, float* out, int ldo, cudaStream_t stream) {
  GVK_CHECK_ARG(a && w && out && M > 0, "gvk_small_matmul: null pointer");
  GVK_CHECK_ARG(ra >= 1 && ra <= 96 && rb >= 1 && rb <= 96 && ra * rb <= 4096, "gvk_small_matmul: ra=%d rb=%d must be in [1,96], ra*rb <= 4096", ra, rb);
  const bool vec = lda % 4 == 0 && ldo % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec && ra == 60 && rb == 20) {
    small_matmul_rows_kernel<60, 20><<<(M + 127) / 128, 128, 0, stream>>>(a, lda, w, M, out, ldo);
  } else if (vec && ra == 96 && rb == 32) {
    small_matmul_rows_kernel<96, 32><<<(M + 127) / 128, 128, 0, stream>>>(a, lda, w, M, out, ldo);
  } else {
    const size_t total = (size_t)M * rb;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 8);
    small_matmul_kernel<<<grid, 256, 0, stream>>>(a, lda, ra, w, rb, M, out, ldo);
  }
  GVK_CHECK_LAUNCH("small_matmul");
  return GVK_OK;
}

// dst[row, s * r + j] = (pattern bit s ? lo : hi)(src[row, j]) for s = 0..2, zero up to `width`:  hi = bf16(x), lo = bf16(x - hi).
// With A rows packed as (hi, lo, hi) and B rows as (hi, hi, lo) a bf16 GEMM over these 3 r columns yields hi*hi + lo*hi + hi*lo, i.e. the
// fp32 product to ~2^-16 relative: this is how the rank-r prompt up-projection rides on the fc2 GEMM as one extra K block.
__global__ void __launch_bounds__(256) split_pack_bf16_kernel(const float* __restrict__ src, int ld_src, int rows, int r, __nv_bfloat16* __restrict__ dst, int ld_dst,
                                                              int width, int pattern) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * width) return;
  const int row = idx / width, c = idx - row * width;
  const int slot = c / r, j = c - slot * r;
  float out = 0.f;
  if (slot < 3) {
    const float x = src[(size_t)row * ld_src + j];
    const float hi = __bfloat162float(__float2bfloat16_rn(x));
    out = (pattern >> slot) & 1 ? x - hi : hi;
  }
  dst[(size_t)row * ld_dst + c] = __float2bfloat16_rn(out);
}
// Same rule, a thread per 8 output columns and one 16-byte store (width % 8 == 0, 16-byte aligned rows): the element-per-thread form spends
// 23 us on the [66 k, 20] -> [66 k, 64] pack of a GAViKO layer (two integer divisions and a 2-byte store per element) for 13 MB of traffic.
__global__ void __launch_bounds__(256) split_pack_bf16_vec_kernel(const float* __restrict__ src, int ld_src, int rows, int r, __nv_bfloat16* __restrict__ dst, int ld_dst,
                                                                  int width, int pattern) {
  const int per_row = width / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * per_row) return;
  const int row = idx / per_row, c0 = (idx - row * per_row) * 8;
  const float* sr = src + (size_t)row * ld_src;
  __align__(16) __nv_bfloat16 o[8];
  int slot = c0 / r, j = c0 - slot * r;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float out = 0.f;
    if (slot < 3) {
      const float x = sr[j];
      const float hi = __bfloat162float(__float2bfloat16_rn(x));
      out = (pattern >> slot) & 1 ? x - hi : hi;
    }
    o[e] = __float2bfloat16_rn(out);
    if (++j == r) { j = 0; ++slot; }
  }
  *reinterpret_cast<uint4*>(dst + (size_t)row * ld_dst + c0) = *reinterpret_cast<const uint4*>(o);
}

int split_pack_bf16(const float* src, int ld_src, int rows, int r, void* dst, int ld_dst, int width, int pattern, cudaStream_t stream) {
  GVK_CHECK_ARG(src && dst && rows > 0 && r > 0 && width >= 3 * r, "gvk_split_pack_bf16: bad argument (rows=%d r=%d width=%d)", rows, r, width);
  if (width % 8 == 0 && ld_dst % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const long long total = (long long)rows * (width / 8);
    split_pack_bf16_vec_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(src, ld_src, rows, r, reinterpret_cast<__nv_bfloat16*>(dst), ld_dst, width, pattern);
    GVK_CHECK_LAUNCH("split_pack_bf16");
    return GVK_OK;
  }
  const long long total = (long long)rows * width;
  split_pack_bf16_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(src, ld_src, rows, r, reinterpret_cast<__nv_bfloat16*>(dst), ld_dst, width, pattern);
  GVK_CHECK_LAUNCH("split_pack_bf16");
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

// ------------------------------------------------------------------------------------------------
// SSF site backward / bias gradients (include/gvk.h: gvk_ssf_bwd)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ssf_bwd_kernel(gvk_ssf_bwd_params p, int rows_per_cta) {
  const int n = blockIdx.y * 256 + threadIdx.x;
  if (n >= p.N) return;
  const int m_begin = blockIdx.x * rows_per_cta, m_end = min(p.M, m_begin + rows_per_cta);
  const T* dy = reinterpret_cast<const T*>(p.dy);
  const T* y = reinterpret_cast<const T*>(p.y);
  T* dx = reinterpret_cast<T*>(p.dx);
  const float sc = p.scale ? p.scale[n] : 1.f, sh = p.shift ? p.shift[n] : 0.f;
  const float inv_sc = 1.0f / sc;
  float ds = 0.f, db = 0.f;
  for (int m = m_begin; m < m_end; ++m) {
    int row = m, r = 0;
    if (p.rows_per_batch > 0) {
      const int b = m / p.rows_per_batch;
      r = m - b * p.rows_per_batch;
      row = b * p.batch_rows + r;
    }
    const float g = ld_as_float(dy + (size_t)row * p.ld_dy + n);
    db += g;
    if (p.scale) {
      float xin = ld_as_float(y + (size_t)row * p.ld_y + n);
      if (p.sub) xin -= p.sub[(size_t)r * p.ld_sub + n];
      xin = (xin - sh) * inv_sc;
      ds = fmaf(g, xin, ds);
      if (dx) st_from_float(dx + (size_t)row * p.ld_dx + n, g * sc);
    }
  }
  if (p.dshift) atomicAdd(p.dshift + n, db);
  if (p.scale && p.dscale) atomicAdd(p.dscale + n, ds);
}

int ssf_bwd(const gvk_ssf_bwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->dy && p->M > 0 && p->N > 0, "gvk_ssf_bwd: bad argument");
  GVK_CHECK_ARG(!p->scale || (p->y && p->shift), "gvk_ssf_bwd: the SSF form needs y and shift");
  GVK_CHECK_ARG(!p->sub || p->rows_per_batch > 0, "gvk_ssf_bwd: sub needs rows_per_batch");
  const int ctas = std::max(1, std::min(sm_count() * 8 / ((p->N + 255) / 256), (p->M + 15) / 16));
  const int rows_per_cta = (p->M + ctas - 1) / ctas;
  dim3 grid((p->M + rows_per_cta - 1) / rows_per_cta, (p->N + 255) / 256);
  if (p->dtype == GVK_F32) ssf_bwd_kernel<float><<<grid, 256, 0, stream>>>(*p, rows_per_cta);
  else ssf_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(*p, rows_per_cta);
  GVK_CHECK_LAUNCH("ssf_bwd");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// elementwise dropout (+ residual)
// ------------------------------------------------------------------------------------------------
// Mask rule (include/gvk.h, gvk_dropout_params): element e = offset + m * N + n is kept iff byte (e % 16) of
// philox4x32-10(counter = (e / 16 low, e / 16 high, 0, 'drop'), key = seed) is < thr = round(256 (1 - drop_p)); kept values are scaled by 256 / thr
// (the rule of the attention-probability dropout, gvk_mhsa_fwd_params).  One Philox call decides 16 elements, so a pass over an [M, 3072] bf16
// activation is bound by HBM, not by the integer pipe (the 4-elements-per-call rule of the rank-r kernels cost ~20 instructions per element).
constexpr uint32_t kDropTag = 0x64726f70u;   // 'drop'

template <typename T>
__device__ __forceinline__ void drop_ld4(const T* p, bool vec, float (&v)[4]) {
  if (vec) {
    if constexpr (sizeof(T) == 4) {
      const float4 q = *reinterpret_cast<const float4*>(p);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      const uint2 q = *reinterpret_cast<const uint2*>(p);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld_as_float<T>(p + u);
  }
}
template <typename T>
__device__ __forceinline__ void drop_st4(T* p, bool vec, const float (&v)[4]) {
  if (vec) {
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
      *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) st_from_float<T>(p + u, v[u]);
  }
}

// 4 consecutive elements starting at column c of row m, decided by the 4 bytes of `word`
template <typename TX, typename TO>
__device__ __forceinline__ void dropout_quad(const gvk_dropout_params& p, uint32_t word, uint32_t thr, float scale, int m, int c, bool vec) {
  float v[4];
  drop_ld4<TX>(reinterpret_cast<const TX*>(p.x) + (size_t)m * p.ldx + c, vec, v);
#pragma unroll
  for (int u = 0; u < 4; ++u) v[u] = ((word >> (8 * u)) & 0xFFu) < thr ? v[u] * scale : 0.f;
  if (p.res) {
    float r[4];
    drop_ld4<float>(p.res + (size_t)m * p.ld_res + c, vec, r);
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] += r[u];
  }
  drop_st4<TO>(reinterpret_cast<TO*>(p.out) + (size_t)m * p.ld_out + c, vec, v);
}

template <typename TX, typename TO, bool WIDE>
__global__ void __launch_bounds__(256) dropout_kernel(gvk_dropout_params p, int vec, uint32_t thr) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  const uint2 key = make_uint2((uint32_t)seed_eff, (uint32_t)(seed_eff >> 32));
  const float scale = 256.0f / (float)thr;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if constexpr (WIDE) {      // N % 16 == 0 and offset % 16 == 0: a thread owns a whole 16-element group: one Philox call, 16-byte accesses
                             // (every lane running its own call is what makes the call cheap per element: a divergent "one lane in four"
                             // issues the same instructions for a quarter of the work)
    const uint32_t n16 = p.N / 16;
    const size_t total = (size_t)p.M * n16;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
      const uint32_t m = (uint32_t)(i / n16), c = (uint32_t)(i - (size_t)m * n16) * 16;
      const uint64_t ctr = (p.offset + (uint64_t)m * p.N + c) >> 4;
      const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, kDropTag), key);
      const uint32_t words[4] = {r.x, r.y, r.z, r.w};
      if (vec == 2) {        // 16-byte aligned rows: bf16 travels as two 8-element accesses, fp32 as four 4-element ones
        float v[16];
        const TX* x = reinterpret_cast<const TX*>(p.x) + (size_t)m * p.ldx + c;
        if constexpr (sizeof(TX) == 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 q = reinterpret_cast<const float4*>(x)[u];
            v[4 * u] = q.x; v[4 * u + 1] = q.y; v[4 * u + 2] = q.z; v[4 * u + 3] = q.w;
          }
        } else {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint4 q = reinterpret_cast<const uint4*>(x)[u];
            const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qq[k]));
              v[8 * u + 2 * k] = f.x; v[8 * u + 2 * k + 1] = f.y;
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = ((words[e >> 2] >> (8 * (e & 3))) & 0xFFu) < thr ? v[e] * scale : 0.f;
        if (p.res) {
          const float4* rs = reinterpret_cast<const float4*>(p.res + (size_t)m * p.ld_res + c);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 q = rs[u];
            v[4 * u] += q.x; v[4 * u + 1] += q.y; v[4 * u + 2] += q.z; v[4 * u + 3] += q.w;
          }
        }
        TO* o = reinterpret_cast<TO*>(p.out) + (size_t)m * p.ld_out + c;
        if constexpr (sizeof(TO) == 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) reinterpret_cast<float4*>(o)[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
        } else {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t qq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * u + 2 * k], v[8 * u + 2 * k + 1]);
              qq[k] = *reinterpret_cast<uint32_t*>(&h);
            }
            reinterpret_cast<uint4*>(o)[u] = make_uint4(qq[0], qq[1], qq[2], qq[3]);
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) dropout_quad<TX, TO>(p, words[u], thr, scale, (int)m, (int)c + 4 * u, vec != 0);
      }
    }
  } else {                   // any N % 4 == 0, offset % 4 == 0: a thread owns 4 elements and takes its word of the group's call
    const uint32_t n4 = p.N / 4;
    const size_t total = (size_t)p.M * n4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
      const uint32_t m = (uint32_t)(i / n4), c = (uint32_t)(i - (size_t)m * n4) * 4;
      const uint64_t e = p.offset + (uint64_t)m * p.N + c, ctr = e >> 4;
      const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, kDropTag), key);
      const uint32_t q = (uint32_t)(e >> 2) & 3u;
      dropout_quad<TX, TO>(p, q == 0 ? r.x : q == 1 ? r.y : q == 2 ? r.z : r.w, thr, scale, (int)m, (int)c, vec != 0);
    }
  }
}

template <typename TX, typename TO>
static void dropout_launch(const gvk_dropout_params* p, int vec, uint32_t thr, cudaStream_t stream) {
  const bool wide = p->N % 16 == 0 && (p->offset & 15) == 0;
  const size_t total = (size_t)p->M * (p->N / (wide ? 16 : 4));
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16));
  if (wide) dropout_kernel<TX, TO, true><<<grid, 256, 0, stream>>>(*p, vec, thr);
  else dropout_kernel<TX, TO, false><<<grid, 256, 0, stream>>>(*p, vec, thr);
}

int dropout(const gvk_dropout_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->x && p->out && p->M > 0 && p->N > 0 && p->N % 4 == 0, "gvk_dropout: bad argument (N %% 4 == 0)");
  GVK_CHECK_ARG(p->drop_p >= 0.f && p->drop_p < 1.f && (p->offset & 3) == 0, "gvk_dropout: drop_p in [0,1), offset %% 4 == 0");
  GVK_CHECK_ARG((p->x_dtype == GVK_F32 || p->x_dtype == GVK_BF16) && (p->out_dtype == GVK_F32 || p->out_dtype == GVK_BF16), "gvk_dropout: dtypes must be GVK_F32 / GVK_BF16");
  const uint32_t thr = (uint32_t)std::min(256, std::max(1, (int)(256.0f * (1.0f - p->drop_p) + 0.5f)));
  auto al = [](const void* q, int dtype, int ld) { return (reinterpret_cast<uintptr_t>(q) & (dtype == GVK_F32 ? 15 : 7)) == 0 && ld % 4 == 0; };
  int vec = al(p->x, p->x_dtype, p->ldx) && al(p->out, p->out_dtype, p->ld_out) && (!p->res || al(p->res, GVK_F32, p->ld_res));
  auto al16 = [](const void* q, int dtype, int ld) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0 && ld % (dtype == GVK_F32 ? 4 : 8) == 0; };
  if (vec && al16(p->x, p->x_dtype, p->ldx) && al16(p->out, p->out_dtype, p->ld_out)) vec = 2;      // 16-byte rows on both sides (res: fp32, already 16-byte)
  if (p->x_dtype == GVK_F32 && p->out_dtype == GVK_F32) dropout_launch<float, float>(p, vec, thr, stream);
  else if (p->x_dtype == GVK_F32) dropout_launch<float, __nv_bfloat16>(p, vec, thr, stream);
  else if (p->out_dtype == GVK_F32) dropout_launch<__nv_bfloat16, float>(p, vec, thr, stream);
  else dropout_launch<__nv_bfloat16, __nv_bfloat16>(p, vec, thr, stream);
  GVK_CHECK_LAUNCH("dropout");
  return GVK_OK;
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = z[i] > 0.f ? dy[i] : 0.f;
}
int relu_bwd(const float* dy, const float* z, float* y, size_t n, cudaStream_t stream) {
  GVK_CHECK_ARG(dy && z && y && n > 0, "gvk_relu_bwd: bad argument");
  relu_bwd_kernel<<<(int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 16), 256, 0, stream>>>(dy, z, y, n);
  GVK_CHECK_LAUNCH("relu_bwd");
  return GVK_OK;
}

__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int M, int dim2) {
  const size_t total = (size_t)M * dim2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / dim2), c = (int)(i - (size_t)m * dim2) * 2;
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + (size_t)m * ldx + c));
    *reinterpret_cast<float2*>(y + (size_t)m * ldy + c) = v;
  }
}

int cast_bf16_f32(const void* x, int ldx, float* y, int ldy, int M, int dim, cudaStream_t stream) {
  GVK_CHECK_ARG(x && y && M > 0 && dim > 0 && dim % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0, "gvk_cast_bf16_f32: bad argument");
  const size_t total = (size_t)M * (dim / 2);
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
  cast_bf16_f32_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, y, ldy, M, dim / 2);
  GVK_CHECK_LAUNCH("cast_bf16_f32");
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
