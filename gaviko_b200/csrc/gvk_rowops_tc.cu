// gvk_rowops_tc.cu — tensor-core forms of the rank-r row kernels (precision = GVK_PREC_TF32 in include/gvk.h).
//
// The exact-fp32 forms in gvk_rowops.cu are bound by FMA issue (M * dim * r FMAs ~ 15 us on 148 SMs for M = 33k, dim = 768, r = 20,
// the same as the HBM time of the one pass over the [M, dim] operand), so they cannot get within 2x of the memory roofline.  Here the
// rank-r products run as mma.sync m16n8k8 (tf32 operands rounded to nearest, fp32 accumulate) on register fragments loaded straight
// from global memory with 16-byte accesses; what is left is the HBM stream.  The k (or n) index inside a 16-wide group is permuted so
// that one float4 per lane feeds two MMAs (both operands use the same permutation, so the product is unchanged).
//
// Used by the bf16 compute mode for LocalSelfAttention / Awakening_Prompt projections and their weight gradients
// (reference model/gaviko.py:149-187, 229-244 and the autograd of those lines).
#include <algorithm>
#include <cstdlib>

#include "gvk_common.cuh"

namespace gvk {

constexpr int kTcWarps = 8;
constexpr int kTcThreads = kTcWarps * 32;

__device__ __forceinline__ float2 tc_drop2(uint64_t seed, uint64_t e, float p, float inv_keep) {
  const uint64_t ctr = e >> 2;
  const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const bool hi = (e & 2) != 0;
  const uint32_t a = hi ? r.z : r.x, b = hi ? r.w : r.y;
  return make_float2(u32_to_unit(a) >= p ? inv_keep : 0.f, u32_to_unit(b) >= p ? inv_keep : 0.f);
}
// mask of 4 consecutive elements starting at e (e % 4 == 0): one Philox call
__device__ __forceinline__ float4 tc_drop4(uint64_t seed, uint64_t e, float p, float inv_keep) {
  const uint64_t ctr = e >> 2;
  const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return make_float4(u32_to_unit(r.x) >= p ? inv_keep : 0.f, u32_to_unit(r.y) >= p ? inv_keep : 0.f, u32_to_unit(r.z) >= p ? inv_keep : 0.f,
                     u32_to_unit(r.w) >= p ? inv_keep : 0.f);
}


// Stage a rank-r panel w(j, c) (strided, see include/gvk.h) into shared memory as dst[j * S + perm(c)] = tf32(scale[c] * w(j, c)) for
// j < RP (rows >= r zero), reading global memory with 16-byte loads in its own linear order, all loads of a thread in flight at once.
// PERM swaps the two low column bits (tc_up); scale may be null.
template <int RP, int S, bool PERM>
__device__ __forceinline__ void tc_stage_panel(float* dst, const float* __restrict__ w, int r, int dim, int w_sj, int w_sc, const float* __restrict__ scale) {
  const int tid = threadIdx.x;
  auto pos = [](int c) { return PERM ? ((c & ~3) | ((c & 1) << 1) | ((c >> 1) & 1)) : c; };
  const bool vec_ok = (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (scale == nullptr || (reinterpret_cast<uintptr_t>(scale) & 15) == 0);
  if (w_sc == 1 && w_sj % 4 == 0 && dim % 4 == 0 && vec_ok) {           // [r, dim] row-major: float4 along c
    const int nv = r * (dim / 4);
    for (int v0 = tid; v0 < nv; v0 += kTcThreads * 4) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = min(v0 + u * kTcThreads, nv - 1);
        const int j = v / (dim / 4), c = (v - j * (dim / 4)) * 4;
        q[u] = *reinterpret_cast<const float4*>(w + (size_t)j * w_sj + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * kTcThreads;
        if (v >= nv) break;
        const int j = v / (dim / 4), c = (v - j * (dim / 4)) * 4;
        float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (scale) sc4 = *reinterpret_cast<const float4*>(scale + c);
        if (PERM) {
          dst[j * S + pos(c)] = tf32_round(q[u].x * sc4.x); dst[j * S + pos(c + 1)] = tf32_round(q[u].y * sc4.y);
          dst[j * S + pos(c + 2)] = tf32_round(q[u].z * sc4.z); dst[j * S + pos(c + 3)] = tf32_round(q[u].w * sc4.w);
        } else {
          *reinterpret_cast<float4*>(dst + j * S + c) = make_float4(tf32_round(q[u].x * sc4.x), tf32_round(q[u].y * sc4.y), tf32_round(q[u].z * sc4.z), tf32_round(q[u].w * sc4.w));
        }
      }
    }
  } else if (w_sj == 1 && w_sc == r && r % 4 == 0 && vec_ok) {            // [dim, r] row-major: float4 along j
    const int nv = dim * (r / 4);
    for (int v0 = tid; v0 < nv; v0 += kTcThreads * 4) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = min(v0 + u * kTcThreads, nv - 1);
        q[u] = *reinterpret_cast<const float4*>(w + (size_t)v * 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * kTcThreads;
        if (v >= nv) break;
        const int c = v / (r / 4), j = (v - c * (r / 4)) * 4;
        const float sc = scale ? scale[c] : 1.f;
        const int pc = pos(c);
        dst[j * S + pc] = tf32_round(q[u].x * sc); dst[(j + 1) * S + pc] = tf32_round(q[u].y * sc);
        dst[(j + 2) * S + pc] = tf32_round(q[u].z * sc); dst[(j + 3) * S + pc] = tf32_round(q[u].w * sc);
      }
    }
  } else {
    for (int idx = tid; idx < r * dim; idx += kTcThreads) {
      int j, c;
      if (w_sc == 1) { j = idx / dim; c = idx - j * dim; } else { c = idx / r; j = idx - c * r; }
      dst[j * S + pos(c)] = tf32_round(w[(size_t)j * w_sj + (size_t)c * w_sc] * (scale ? scale[c] : 1.f));
    }
  }
  for (int idx = r * dim + tid; idx < RP * dim; idx += kTcThreads) {       // zero rows r .. RP-1 (positions are a permutation: any order)
    const int j = idx / dim, c = idx - j * dim;
    dst[j * S + c] = 0.f;
  }
}

static inline int tc_grid(int tiles, int per_sm) { return std::max(1, std::min((tiles + kTcWarps - 1) / kTcWarps, sm_count() * per_sm)); }

// =================================================================================================
// down projection:  z = act(LN?(drop?(x)) W^T + b)  (+ pre-activation, + chained z2 = z W2^T)
// A warp owns 16 rows; lane (g, t) streams rows g and g+8 as float4 at columns 16 kk + 4 t.
// LayerNorm is folded algebraically: LN(x) . w_j = rstd (x . (gamma * w_j) - mean sum(gamma * w_j)) + beta . w_j, with mean / rstd taken
// from the same registers in the same pass.
// =================================================================================================
template <int NITER, int NT>
__global__ void __launch_bounds__(kTcThreads, NT <= 4 ? 2 : 1) tc_down_kernel(gvk_rowproj_down_params p) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  constexpr int dim = NITER * 64, S = dim + 16, KG = dim / 16, RP = NT * 8;
  constexpr int S2 = 40;  // chained-projection panel row stride (floats): conflict-free float2 fragment loads for RP <= 32
  constexpr int U = (KG % 8 == 0) ? 8 : 4;   // 16-column groups loaded per batch (2 U float4 in flight per lane); KG % U == 0
  static_assert(KG % U == 0, "column groups must tile the row");
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;            // [RP][S]  tf32(gamma * W), rows >= r zero
  float* s_cs = sW + RP * S;   // [RP]     sum_c sW[j][c]
  float* s_b = s_cs + RP;      // [RP]     b_j + beta . w_j
  float* sW2 = s_b + RP;       // [r2pad][S2] tf32(W2), zero padded
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const bool has_ln = p.ln_gamma != nullptr;
  const int ntiles = (p.M + 15) / 16;
  // (An L2 prefetch of the warp's next tile was tried here and in tc_up / layernorm_bwd: it cost 20 % — 41 % more DRAM reads and LSU
  // queue stalls — because the 16 float4 loads a lane keeps in flight already cover the memory latency.)
  tc_stage_panel<RP, S, false>(sW, p.w, p.r, dim, p.w_sj, p.w_sc, p.ln_gamma);
  const int r2pad = p.w2 ? (p.r2 + 7) / 8 * 8 : 0;
  for (int idx = tid; idx < r2pad * S2; idx += kTcThreads) {
    const int k = idx / S2, j = idx - k * S2;
    sW2[idx] = (k < p.r2 && j < p.r) ? tf32_round(p.w2[k * p.r + j]) : 0.f;
  }
  __syncthreads();
  for (int j = warp; j < RP; j += kTcWarps) {
    float cs = 0.f, bb = 0.f;
    if (has_ln && j < p.r) {
      for (int c = lane; c < dim; c += 32) {
        cs += sW[j * S + c];
        bb = fmaf(p.ln_beta[c], p.w[(size_t)j * p.w_sj + (size_t)c * p.w_sc], bb);
      }
      cs = warp_sum(cs);
      bb = warp_sum(bb);
    }
    if (lane == 0) {
      s_cs[j] = cs;
      s_b[j] = bb + ((p.bias && j < p.r) ? p.bias[j] : 0.f);
    }
  }
  __syncthreads();

  const float inv_dim = 1.0f / dim;
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  for (int tile = blockIdx.x * kTcWarps + warp; tile < ntiles; tile += gridDim.x * kTcWarps) {
    const int rA = tile * 16 + g, rB = rA + 8;
    const size_t cA = (size_t)min(rA, p.M - 1), cB = (size_t)min(rB, p.M - 1);
    const float* xa = p.x + cA * p.ldx + 4 * t;
    const float* xb = p.x + cB * p.ldx + 4 * t;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    float sA = 0.f, qA = 0.f, sB = 0.f, qB = 0.f;
#pragma unroll 1
    for (int k0 = 0; k0 < KG; k0 += U) {
      float4 va[U], vb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        va[u] = *reinterpret_cast<const float4*>(xa + 16 * (k0 + u));
        vb[u] = *reinterpret_cast<const float4*>(xb + 16 * (k0 + u));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kk = k0 + u;
        if (p.drop_p > 0.f) {
          const float4 ma = tc_drop4(seed_eff, p.offset + cA * dim + 16 * kk + 4 * t, p.drop_p, inv_keep);
          const float4 mb = tc_drop4(seed_eff, p.offset + cB * dim + 16 * kk + 4 * t, p.drop_p, inv_keep);
          va[u].x *= ma.x; va[u].y *= ma.y; va[u].z *= ma.z; va[u].w *= ma.w;
          vb[u].x *= mb.x; vb[u].y *= mb.y; vb[u].z *= mb.z; vb[u].w *= mb.w;
        }
        if (has_ln) {
          sA += (va[u].x + va[u].y) + (va[u].z + va[u].w);
          qA = fmaf(va[u].x, va[u].x, fmaf(va[u].y, va[u].y, fmaf(va[u].z, va[u].z, fmaf(va[u].w, va[u].w, qA))));
          sB += (vb[u].x + vb[u].y) + (vb[u].z + vb[u].w);
          qB = fmaf(vb[u].x, vb[u].x, fmaf(vb[u].y, vb[u].y, fmaf(vb[u].z, vb[u].z, fmaf(vb[u].w, vb[u].w, qB))));
        }
        const uint32_t ax = f2tf32(va[u].x), ay = f2tf32(va[u].y), az = f2tf32(va[u].z), aw = f2tf32(va[u].w);
        const uint32_t bx = f2tf32(vb[u].x), by = f2tf32(vb[u].y), bz = f2tf32(vb[u].z), bw = f2tf32(vb[u].w);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(sW + (8 * j + g) * S + 16 * kk + 4 * t);
          mma_tf32(acc[j], ax, bx, ay, by, __float_as_uint(w.x), __float_as_uint(w.y));
          mma_tf32(acc[j], az, bz, aw, bw, __float_as_uint(w.z), __float_as_uint(w.w));
        }
      }
    }
    float meanA = 0.f, rstdA = 1.f, meanB = 0.f, rstdB = 1.f;
    if (has_ln) {
      sA += __shfl_xor_sync(0xffffffffu, sA, 1); sA += __shfl_xor_sync(0xffffffffu, sA, 2);
      qA += __shfl_xor_sync(0xffffffffu, qA, 1); qA += __shfl_xor_sync(0xffffffffu, qA, 2);
      sB += __shfl_xor_sync(0xffffffffu, sB, 1); sB += __shfl_xor_sync(0xffffffffu, sB, 2);
      qB += __shfl_xor_sync(0xffffffffu, qB, 1); qB += __shfl_xor_sync(0xffffffffu, qB, 2);
      meanA = sA * inv_dim; meanB = sB * inv_dim;
      rstdA = rsqrtf(fmaxf(qA * inv_dim - meanA * meanA, 0.f) + p.eps);
      rstdB = rsqrtf(fmaxf(qB * inv_dim - meanB * meanB, 0.f) + p.eps);
      if (t == 0) {
        if (rA < p.M) { if (p.mean) p.mean[rA] = meanA; if (p.rstd) p.rstd[rA] = rstdA; }
        if (rB < p.M) { if (p.mean) p.mean[rB] = meanB; if (p.rstd) p.rstd[rB] = rstdB; }
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 8 * j + 2 * t + e;
        float zA = acc[j][e], zB = acc[j][2 + e];
        if (has_ln) {
          zA = rstdA * (zA - meanA * s_cs[n]);
          zB = rstdB * (zB - meanB * s_cs[n]);
        }
        zA += s_b[n];
        zB += s_b[n];
        if (n < p.r && p.pre) {
          if (rA < p.M) p.pre[(size_t)rA * p.ldz + n] = zA;
          if (rB < p.M) p.pre[(size_t)rB * p.ldz + n] = zB;
        }
        if (p.act == GVK_ROWACT_QUICKGELU) {
          zA = quick_gelu(zA); zB = quick_gelu(zB);
        } else if (p.act == GVK_ROWACT_RELU) {
          zA = fmaxf(zA, 0.f); zB = fmaxf(zB, 0.f);
        }
        if (n >= p.r) zA = zB = 0.f;
        if (n < p.r) {
          if (rA < p.M) p.z[(size_t)rA * p.ldz + n] = zA;
          if (rB < p.M) p.z[(size_t)rB * p.ldz + n] = zB;
        }
        acc[j][e] = zA;
        acc[j][2 + e] = zB;
      }
    }
    if (p.w2) {
      // z2 = z W2^T: the C fragments of z are exactly the A fragments of the next MMA (k slot t <-> latent 8j+2t, slot t+4 <-> 8j+2t+1)
      for (int jj = 0; jj < r2pad / 8; ++jj) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float2 w = *reinterpret_cast<const float2*>(sW2 + (8 * jj + g) * S2 + 8 * j + 2 * t);
          mma_tf32(d, f2tf32(acc[j][0]), f2tf32(acc[j][2]), f2tf32(acc[j][1]), f2tf32(acc[j][3]), __float_as_uint(w.x), __float_as_uint(w.y));
        }
        const int n2 = 8 * jj + 2 * t;
        if (n2 + 1 < p.r2) {
          if (rA < p.M) *reinterpret_cast<float2*>(p.z2 + (size_t)rA * p.ldz2 + n2) = make_float2(d[0], d[1]);
          if (rB < p.M) *reinterpret_cast<float2*>(p.z2 + (size_t)rB * p.ldz2 + n2) = make_float2(d[2], d[3]);
        } else if (n2 < p.r2) {
          if (rA < p.M) p.z2[(size_t)rA * p.ldz2 + n2] = d[0];
          if (rB < p.M) p.z2[(size_t)rB * p.ldz2 + n2] = d[2];
        }
      }
    }
  }
}

template <int NITER, int NT>
static int tc_down_launch(const gvk_rowproj_down_params* p, cudaStream_t stream) {
  constexpr int dim = NITER * 64;
  const int r2pad = p->w2 ? (p->r2 + 7) / 8 * 8 : 0;
  const size_t smem = ((size_t)NT * 8 * (dim + 16) + 2 * NT * 8 + (size_t)r2pad * 40) * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("gvk_rowproj_down(tf32): r=%d at dim=%d needs %zu B of shared memory", p->r, p->dim, smem);
    return GVK_ERR_UNSUPPORTED;
  }
  static size_t configured = 0;
  if (smem > configured) {
    int st = cuda_status(cudaFuncSetAttribute(tc_down_kernel<NITER, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tc_down smem");
    if (st != GVK_OK) return st;
    configured = smem;
  }
  const int per_sm = (NT <= 4 && smem <= 110 * 1024) ? 2 : 1;
  tc_down_kernel<NITER, NT><<<tc_grid((p->M + 15) / 16, per_sm), kTcThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("rowproj_down_tc");
  return GVK_OK;
}

int rowproj_down_tc(const gvk_rowproj_down_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p->ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0, "gvk_rowproj_down(tf32): x must be 16-byte aligned with ldx %% 4 == 0");
  GVK_CHECK_ARG(!p->w2 || (p->r <= 32 && p->ldz2 % 2 == 0 && (reinterpret_cast<uintptr_t>(p->z2) & 7) == 0), "gvk_rowproj_down(tf32): chained projection needs r <= 32 and an 8-byte aligned z2");
#define GVK_TC_DOWN(NITER)                                              \
  if (p->r <= 8) return tc_down_launch<NITER, 1>(p, stream);            \
  if (p->r <= 24) return tc_down_launch<NITER, 3>(p, stream);           \
  if (p->r <= 32) return tc_down_launch<NITER, 4>(p, stream);           \
  return tc_down_launch<NITER, 8>(p, stream);
  GVK_CHECK_ARG(p->r <= 64, "gvk_rowproj_down(tf32): r=%d must be <= 64", p->r);
  switch (p->dim / 64) {
    case 3: GVK_TC_DOWN(3)
    case 6: GVK_TC_DOWN(6)
    case 12: GVK_TC_DOWN(12)
    case 16: GVK_TC_DOWN(16)
    default:
      set_last_error("row kernels support dim in {192, 384, 768, 1024}, got %d", p->dim);
      return GVK_ERR_UNSUPPORTED;
  }
#undef GVK_TC_DOWN
}

// =================================================================================================
// up projection:  out = res + drop(c W + b)  (+ bf16 copy).  W is staged with the two low column bits of every 16-column group swapped,
// which makes lane (g, t) own the four consecutive output columns 16 kk + 4 t .. + 3 of rows g and g+8 (float4 stores).
// =================================================================================================
template <int NITER, int KS>
__global__ void __launch_bounds__(kTcThreads, KS <= 4 ? 2 : 1) tc_up_kernel(gvk_rowproj_up_params p) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  constexpr int dim = NITER * 64, S = dim + 8, KG = dim / 16, RP = KS * 8;
  constexpr int U = 4;
  static_assert(KG % U == 0, "column groups must tile the row");
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;           // [RP][S]
  float* s_bias = sW + RP * S;  // [dim]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (p.M + 15) / 16;
  tc_stage_panel<RP, S, true>(sW, p.w, p.r, dim, p.w_sj, p.w_sc, nullptr);
  for (int c = tid; c < dim; c += kTcThreads) s_bias[c] = p.bias ? p.bias[c] : 0.f;
  __syncthreads();
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  for (int tile = blockIdx.x * kTcWarps + warp; tile < ntiles; tile += gridDim.x * kTcWarps) {
    const int rA = tile * 16 + g, rB = rA + 8;
    const size_t cA = (size_t)min(rA, p.M - 1), cB = (size_t)min(rB, p.M - 1);
    uint32_t a[KS][4];
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int k0 = 8 * s + t, k1 = k0 + 4;
      a[s][0] = f2tf32(k0 < p.r ? p.c[cA * p.ldc + k0] : 0.f);
      a[s][1] = f2tf32(k0 < p.r ? p.c[cB * p.ldc + k0] : 0.f);
      a[s][2] = f2tf32(k1 < p.r ? p.c[cA * p.ldc + k1] : 0.f);
      a[s][3] = f2tf32(k1 < p.r ? p.c[cB * p.ldc + k1] : 0.f);
    }
#pragma unroll 1
    for (int k0 = 0; k0 < KG; k0 += U) {
      float4 ra[U], rb[U];
      if (p.res) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          ra[u] = *reinterpret_cast<const float4*>(p.res + cA * p.ld_res + 16 * (k0 + u) + 4 * t);
          rb[u] = *reinterpret_cast<const float4*>(p.res + cB * p.ld_res + 16 * (k0 + u) + 4 * t);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int col = 16 * (k0 + u) + 4 * t;
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const float2 b0 = *reinterpret_cast<const float2*>(sW + (8 * s + t) * S + 16 * (k0 + u) + 2 * g);
          const float2 b1 = *reinterpret_cast<const float2*>(sW + (8 * s + t + 4) * S + 16 * (k0 + u) + 2 * g);
          mma_tf32(d0, a[s][0], a[s][1], a[s][2], a[s][3], __float_as_uint(b0.x), __float_as_uint(b1.x));
          mma_tf32(d1, a[s][0], a[s][1], a[s][2], a[s][3], __float_as_uint(b0.y), __float_as_uint(b1.y));
        }
        const float4 bias = *reinterpret_cast<const float4*>(s_bias + col);
        float4 vA = make_float4(d0[0] + bias.x, d0[1] + bias.y, d1[0] + bias.z, d1[1] + bias.w);
        float4 vB = make_float4(d0[2] + bias.x, d0[3] + bias.y, d1[2] + bias.z, d1[3] + bias.w);
        if (p.drop_p > 0.f) {
          const float4 ma = tc_drop4(seed_eff, p.offset + cA * dim + col, p.drop_p, inv_keep);
          const float4 mb = tc_drop4(seed_eff, p.offset + cB * dim + col, p.drop_p, inv_keep);
          vA.x *= ma.x; vA.y *= ma.y; vA.z *= ma.z; vA.w *= ma.w;
          vB.x *= mb.x; vB.y *= mb.y; vB.z *= mb.z; vB.w *= mb.w;
        }
        if (p.res) {
          vA.x += ra[u].x; vA.y += ra[u].y; vA.z += ra[u].z; vA.w += ra[u].w;
          vB.x += rb[u].x; vB.y += rb[u].y; vB.z += rb[u].z; vB.w += rb[u].w;
        }
        if (rA < p.M) {
          *reinterpret_cast<float4*>(p.out + (size_t)rA * p.ld_out + col) = vA;
          if (p.out_lp) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(vA.x, vA.y), hi = __floats2bfloat162_rn(vA.z, vA.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out_lp) + (size_t)rA * p.ld_out_lp + col) =
                make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
          }
        }
        if (rB < p.M) {
          *reinterpret_cast<float4*>(p.out + (size_t)rB * p.ld_out + col) = vB;
          if (p.out_lp) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(vB.x, vB.y), hi = __floats2bfloat162_rn(vB.z, vB.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out_lp) + (size_t)rB * p.ld_out_lp + col) =
                make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
          }
        }
      }
    }
  }
}

template <int NITER, int KS>
static int tc_up_launch(const gvk_rowproj_up_params* p, cudaStream_t stream) {
  constexpr int dim = NITER * 64;
  const size_t smem = ((size_t)KS * 8 * (dim + 8) + dim) * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("gvk_rowproj_up(tf32): r=%d at dim=%d needs %zu B of shared memory", p->r, p->dim, smem);
    return GVK_ERR_UNSUPPORTED;
  }
  static size_t configured = 0;
  if (smem > configured) {
    int st = cuda_status(cudaFuncSetAttribute(tc_up_kernel<NITER, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tc_up smem");
    if (st != GVK_OK) return st;
    configured = smem;
  }
  const int per_sm = (KS <= 4 && smem <= 110 * 1024) ? 2 : 1;
  tc_up_kernel<NITER, KS><<<tc_grid((p->M + 15) / 16, per_sm), kTcThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("rowproj_up_tc");
  return GVK_OK;
}

int rowproj_up_tc(const gvk_rowproj_up_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p->ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0 && (!p->res || (p->ld_res % 4 == 0 && (reinterpret_cast<uintptr_t>(p->res) & 15) == 0)) &&
                    (!p->out_lp || (p->ld_out_lp % 4 == 0 && (reinterpret_cast<uintptr_t>(p->out_lp) & 7) == 0)),
                "gvk_rowproj_up(tf32): out / res must be 16-byte aligned with leading dimensions %% 4 == 0");
  GVK_CHECK_ARG(p->r <= 64, "gvk_rowproj_up(tf32): r=%d must be <= 64", p->r);
#define GVK_TC_UP(NITER)                                             \
  if (p->r <= 8) return tc_up_launch<NITER, 1>(p, stream);           \
  if (p->r <= 24) return tc_up_launch<NITER, 3>(p, stream);          \
  if (p->r <= 32) return tc_up_launch<NITER, 4>(p, stream);          \
  return tc_up_launch<NITER, 8>(p, stream);
  switch (p->dim / 64) {
    case 3: GVK_TC_UP(3)
    case 6: GVK_TC_UP(6)
    case 12: GVK_TC_UP(12)
    case 16: GVK_TC_UP(16)
    default:
      set_last_error("row kernels support dim in {192, 384, 768, 1024}, got %d", p->dim);
      return GVK_ERR_UNSUPPORTED;
  }
#undef GVK_TC_UP
}

// =================================================================================================
// weight gradient:  dw(j, c) += sum_m a[m, j] f(x[m, c])  as  D[j, c] += A^T[j, m] X[m, c]  with the row index m as the MMA k dimension.
// A warp owns 32 NG columns (NG = 3: 96, NG = 4: 128) and all 32 (padded) latents; lane (g, t) loads x[m0 + t][.. 4 g ..] and
// x[m0 + t + 4][.. 4 g ..] as float4 (n slot g of MMA i <-> column 32 G + 4 g + i).  Per-CTA partials go to the workspace and are
// summed by skinny_wgrad_reduce (deterministic, same layout as the fp32 kernel).
// =================================================================================================
// The operand stream is software-pipelined: every warp copies its own 16-row x (32 NG)-column slice (plus the 16 latent rows and LN
// statistics) with cp.async into a private ring of shared-memory stages, two stages (NG = 3) ahead of the MMAs, so ~100 KB per SM are in
// flight at any time and no warp ever waits on a global load it has just issued.  Rows are padded to a stride of 32 bytes mod 128 so
// that the float4 fragment reads (rows t / t + 4, columns 4 g) are bank-conflict free.
template <int NG>
struct WgPipe {
  static constexpr int kCols = 32 * NG;
  static constexpr int kRS = kCols + 8;            // x row stride (floats)
  static constexpr int kAS = 40;                   // latent row stride: 8 t + g hits 32 distinct banks
  static constexpr int kRows = 16;                 // rows per stage = two 8-row MMA k steps
  static constexpr int kStages = NG == 3 ? 3 : 2;
  static constexpr int kStageFloats = kRows * kRS + kRows * kAS + 2 * kRows;
  static constexpr size_t kSmem = (size_t)kTcWarps * kStages * kStageFloats * sizeof(float);
};

template <int NG>
__global__ void __launch_bounds__(kTcThreads, 1) tc_wgrad_kernel(gvk_skinny_wgrad_params p, int rows_per_cta, int nwarps) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  using L = WgPipe<NG>;
  extern __shared__ __align__(16) float wg_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  if (warp >= nwarps) return;
  const int m_begin = blockIdx.x * rows_per_cta;
  const int m_end = min(p.M, m_begin + rows_per_cta);
  const int cbase = warp * 32 * NG;
  float* ring = wg_smem + (size_t)warp * L::kStages * L::kStageFloats;
  const bool has_ln = p.ln_gamma != nullptr;
  // The column sum of f(x) (bias gradients) rides on the MMA: latent slot r (a padding slot when r < 32) holds 1.0 for every row, so
  // accumulator row r is the column sum.  Only r == 32 needs explicit adds.
  const bool cs_mma = p.dx_colsum != nullptr && p.r < 32, cs_add = p.dx_colsum != nullptr && p.r == 32;
  // latent slots >= r are zero (never written by the copies); slot r is the ones column
  for (int i = lane; i < L::kStages * L::kRows * L::kAS; i += 32) {
    const int st = i / (L::kRows * L::kAS), rem = i - st * (L::kRows * L::kAS);
    ring[st * L::kStageFloats + L::kRows * L::kRS + rem] = (cs_mma && rem % L::kAS == p.r) ? 1.f : 0.f;
  }
  __syncwarp();
  // per-lane copy plan, fixed for the whole kernel: chunk k of a stage is 16 bytes at (row_k, 4 c4_k)
  constexpr int KX = L::kRows * 8 * NG / 32;
  int xsrc[KX], xdst[KX];
#pragma unroll
  for (int k = 0; k < KX; ++k) {
    const int idx = lane + 32 * k, row = idx / (8 * NG), c4 = idx - row * (8 * NG);
    xsrc[k] = row * p.ldx + c4 * 4;
    xdst[k] = row * L::kRS + c4 * 4;
  }
  const int a_chunks = p.r / 4;
  int asrc[3], adst[3];        // 16 rows x r / 4 <= 128 chunks = 4 per lane; r <= 24 needs 3, r in (24, 32] a fourth handled below
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int idx = lane + 32 * k, row = idx / a_chunks, c4 = idx - row * a_chunks;
    asrc[k] = idx < L::kRows * a_chunks ? row * p.lda + c4 * 4 : -1;
    adst[k] = L::kRows * L::kRS + row * L::kAS + c4 * 4;
  }
  auto issue = [&](int step) {
    const int mb = m_begin + step * L::kRows;
    if (mb < m_end) {
      float* st = ring + (step % L::kStages) * L::kStageFloats;
      const float* xs = p.x + (size_t)mb * p.ldx + cbase;
      const float* as_ = p.a + (size_t)mb * p.lda;
      if (mb + L::kRows <= m_end) {
#pragma unroll
        for (int k = 0; k < KX; ++k) cp_async16(st + xdst[k], xs + xsrc[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (asrc[k] >= 0) cp_async16(st + adst[k], as_ + asrc[k]);
        for (int idx = lane + 96; idx < L::kRows * a_chunks; idx += 32) {
          const int row = idx / a_chunks, c4 = idx - row * a_chunks;
          cp_async16(st + L::kRows * L::kRS + row * L::kAS + c4 * 4, as_ + row * p.lda + c4 * 4);
        }
      } else {                                     // last, partial stage of this CTA: rows >= m_end are zero-filled
        for (int idx = lane; idx < L::kRows * 8 * NG; idx += 32) {
          const int row = idx / (8 * NG), c4 = idx - row * (8 * NG);
          const bool ok = mb + row < m_end;
          cp_async16_zfill(st + row * L::kRS + c4 * 4, xs + (ok ? row * p.ldx : 0) + c4 * 4, ok);
        }
        for (int idx = lane; idx < L::kRows * a_chunks; idx += 32) {
          const int row = idx / a_chunks, c4 = idx - row * a_chunks;
          const bool ok = mb + row < m_end;
          cp_async16_zfill(st + L::kRows * L::kRS + row * L::kAS + c4 * 4, as_ + (ok ? row * p.lda : 0) + c4 * 4, ok);
        }
      }
      if (has_ln) {
        const int row = min(mb + (lane & 15), m_end - 1);
        cp_async4(st + L::kRows * L::kRS + L::kRows * L::kAS + lane, (lane < 16 ? p.mean : p.rstd) + row);
      }
    }
    cp_async_commit();
  };
  float acc[NG][4][2][4];
  float cs[NG][4];
#pragma unroll
  for (int G = 0; G < NG; ++G)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cs[G][i] = 0.f;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) acc[G][i][mt][0] = acc[G][i][mt][1] = acc[G][i][mt][2] = acc[G][i][mt][3] = 0.f;
    }
  float as[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  const int nsteps = (m_end - m_begin + L::kRows - 1) / L::kRows;
#pragma unroll
  for (int s0 = 0; s0 < L::kStages - 1; ++s0) issue(s0);
#pragma unroll 1
  for (int step = 0; step < nsteps; ++step) {
    issue(step + L::kStages - 1);
    cp_async_wait<L::kStages - 1>();
    __syncwarp();
    const int mb = m_begin + step * L::kRows;
    float* st = ring + (step % L::kStages) * L::kStageFloats;
    float* sa = st + L::kRows * L::kRS;
    const float* sstat = sa + L::kRows * L::kAS;
    if (cs_mma && mb + L::kRows > m_end) {         // partial stage: no ones for the zero-filled rows (f(0) != 0 under LayerNorm)
      if (lane < L::kRows && mb + lane >= m_end) sa[lane * L::kAS + p.r] = 0.f;
      __syncwarp();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int l0 = 8 * h + t, l1 = l0 + 4;                     // rows of this lane inside the stage
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float f0 = sa[l0 * L::kAS + 16 * mt + g], f1 = sa[l0 * L::kAS + 16 * mt + g + 8];
        const float f2 = sa[l1 * L::kAS + 16 * mt + g], f3 = sa[l1 * L::kAS + 16 * mt + g + 8];
        as[mt][0] += f0 + f2;
        as[mt][1] += f1 + f3;
        a[mt][0] = tf32_bits(f0); a[mt][1] = tf32_bits(f1); a[mt][2] = tf32_bits(f2); a[mt][3] = tf32_bits(f3);
      }
      float mu0 = 0.f, rs0 = 1.f, mu1 = 0.f, rs1 = 1.f;
      if (has_ln) {
        mu0 = sstat[l0]; rs0 = sstat[16 + l0]; mu1 = sstat[l1]; rs1 = sstat[16 + l1];
      }
#pragma unroll
      for (int G = 0; G < NG; ++G) {
        const int col = cbase + 32 * G + 4 * g;
        const float4 x0 = *reinterpret_cast<const float4*>(st + l0 * L::kRS + 32 * G + 4 * g);
        const float4 x1 = *reinterpret_cast<const float4*>(st + l1 * L::kRS + 32 * G + 4 * g);
        float y0[4] = {x0.x, x0.y, x0.z, x0.w}, y1[4] = {x1.x, x1.y, x1.z, x1.w};
        if (p.drop_p > 0.f) {
          const size_t c0 = (size_t)min(mb + l0, m_end - 1), c1 = (size_t)min(mb + l1, m_end - 1);
          const float4 ma = tc_drop4(seed_eff, p.offset + c0 * p.dim + col, p.drop_p, inv_keep);
          const float4 mb4 = tc_drop4(seed_eff, p.offset + c1 * p.dim + col, p.drop_p, inv_keep);
          y0[0] *= ma.x; y0[1] *= ma.y; y0[2] *= ma.z; y0[3] *= ma.w;
          y1[0] *= mb4.x; y1[1] *= mb4.y; y1[2] *= mb4.z; y1[3] *= mb4.w;
        }
        if (has_ln) {
          const float4 gm = *reinterpret_cast<const float4*>(p.ln_gamma + col), bt = *reinterpret_cast<const float4*>(p.ln_beta + col);
          const float gg[4] = {gm.x, gm.y, gm.z, gm.w}, bb[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            y0[i] = (y0[i] - mu0) * rs0 * gg[i] + bb[i];
            y1[i] = (y1[i] - mu1) * rs1 * gg[i] + bb[i];
          }
        }
        if (cs_add) {
          const bool v0 = mb + l0 < m_end, v1 = mb + l1 < m_end;
#pragma unroll
          for (int i = 0; i < 4; ++i) cs[G][i] += (v0 ? y0[i] : 0.f) + (v1 ? y1[i] : 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t b0 = tf32_bits(y0[i]), b1 = tf32_bits(y1[i]);
          mma_tf32(acc[G][i][0], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
          mma_tf32(acc[G][i][1], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
        }
      }
    }
    __syncwarp();   // every lane is done with this stage before the next iteration's copies overwrite the oldest one
  }
  float* ws_dw = p.ws + (size_t)blockIdx.x * p.r * p.dim;
  float* ws_dx = p.ws + (size_t)gridDim.x * p.r * p.dim + (size_t)blockIdx.x * p.dim;
  float* ws_da = p.ws + (size_t)gridDim.x * (p.r + 1) * p.dim + (size_t)blockIdx.x * 32;
  // acc[G][i][mt][k]: k = 0 (j, slot 2t) 1 (j, slot 2t+1) 2 (j+8, slot 2t) 3 (j+8, slot 2t+1), j = 16 mt + g, column = 32 G + 4 slot + i
#pragma unroll
  for (int G = 0; G < NG; ++G)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = 16 * mt + g + 8 * (k >> 1);
        const int col = cbase + 32 * G + 4 * (2 * t + (k & 1));
        if (j < p.r) *reinterpret_cast<float4*>(ws_dw + (size_t)j * p.dim + col) = make_float4(acc[G][0][mt][k], acc[G][1][mt][k], acc[G][2][mt][k], acc[G][3][mt][k]);
      }
  if (cs_mma) {   // accumulator row r is the column sum (ones slot of the latent operand)
#pragma unroll
    for (int G = 0; G < NG; ++G)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (16 * mt + g + 8 * (k >> 1) == p.r)
            *reinterpret_cast<float4*>(ws_dx + cbase + 32 * G + 4 * (2 * t + (k & 1))) = make_float4(acc[G][0][mt][k], acc[G][1][mt][k], acc[G][2][mt][k], acc[G][3][mt][k]);
  } else {
#pragma unroll
    for (int G = 0; G < NG; ++G) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cs[G][i] += __shfl_xor_sync(0xffffffffu, cs[G][i], 1);
        cs[G][i] += __shfl_xor_sync(0xffffffffu, cs[G][i], 2);
      }
      if (t == 0) *reinterpret_cast<float4*>(ws_dx + cbase + 32 * G + 4 * g) = make_float4(cs[G][0], cs[G][1], cs[G][2], cs[G][3]);
    }
  }
  if (warp == 0) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = as[mt][h];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (t == 0) ws_da[16 * mt + 8 * h + g] = v;
      }
  }
}

void skinny_wgrad_plan(int M, int* ctas, int* rows_per_cta);
void skinny_wgrad_launch_reduce(const gvk_skinny_wgrad_params* p, int ncta, cudaStream_t stream);

int skinny_wgrad_tc(const gvk_skinny_wgrad_params* p, cudaStream_t stream) {
  int grid, rows_per_cta;
  skinny_wgrad_plan(p->M, &grid, &rows_per_cta);
  {   // one CTA per SM (the ring takes all of its shared memory): a single wave, measured 50 us vs 60 us with two CTAs per SM in sequence
    const int want = std::max(1, std::min(grid, sm_count()));
    rows_per_cta = ((p->M + want - 1) / want + 15) / 16 * 16;
    grid = (p->M + rows_per_cta - 1) / rows_per_cta;
  }
  if (p->dim % 128 == 0 && p->dim % 96 != 0) {
    static const int attr4 = cudaFuncSetAttribute(tc_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgPipe<4>::kSmem);
    if (attr4 != cudaSuccess) return cuda_status((cudaError_t)attr4, "skinny_wgrad_tc (smem attribute)");
    tc_wgrad_kernel<4><<<grid, kTcThreads, WgPipe<4>::kSmem, stream>>>(*p, rows_per_cta, p->dim / 128);
  } else {
    static const int attr3 = cudaFuncSetAttribute(tc_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgPipe<3>::kSmem);
    if (attr3 != cudaSuccess) return cuda_status((cudaError_t)attr3, "skinny_wgrad_tc (smem attribute)");
    tc_wgrad_kernel<3><<<grid, kTcThreads, WgPipe<3>::kSmem, stream>>>(*p, rows_per_cta, p->dim / 96);
  }
  GVK_CHECK_LAUNCH("skinny_wgrad_tc");
  skinny_wgrad_launch_reduce(p, grid, stream);
  GVK_CHECK_LAUNCH("skinny_wgrad_reduce");
  return GVK_OK;
}

bool skinny_wgrad_tc_supported(const gvk_skinny_wgrad_params* p) {
  const bool dim_ok = (p->dim % 96 == 0 && p->dim / 96 <= kTcWarps) || (p->dim % 128 == 0 && p->dim / 128 <= kTcWarps);
  return dim_ok && p->r <= 32 && p->r % 4 == 0 && p->ldx % 4 == 0 && p->lda % 4 == 0 && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p->a) & 15) == 0;
}


// =================================================================================================
// LayerNorm backward with a rank-r term, the rank-r product on the tensor cores:
//   MODE 0:  dx = dres + LN'(dy) + az @ aw        dy dense bf16 (the d(g_mid) pass: model/gaviko.py:155 next to :304's LayerNorm)
//   MODE 1:  dx = dres + LN'(dz @ w)              + dgamma / dbeta (LocalSelfAttention.norm -> proj_down, model/gaviko.py:229-231)
// The exact-fp32 kernel (gvk_rowops.cu) re-reads the [r, dim] panel from shared memory for every pair of rows (61 KB per pair at r = 20,
// dim = 768: 232 us against 155 us without the term at M = 66 k).  Here a CTA owns 16 rows per step and warp w owns columns
// [w dim/8, (w+1) dim/8) of all 16 in the C-fragment layout of tc_up (lane (g, t): rows g / g+8, four consecutive columns 16 kk + 4 t), so
// the panel is read once per 16 rows as MMA B fragments and the row statistics need one shared-memory exchange between the 8 warps
// (double-buffered: one __syncthreads per step).
// One CTA per SM.  The three row streams (x, dy, dres) of the NEXT step travel by cp.async into thread-private shared-memory slots while
// the current step is computed: a thread re-fills a slot right after it has read it, so ~120 KB per SM are in flight all the time and no
// warp waits on a load it has just issued.  (Measured at M = 66 112, dim 768, r 20: exact kernel 238 us; this layout with two CTAs per SM
// and plain loads 227 us; one CTA with the next step's x / dy prefetched into registers 182 us; this form 156 us; the same with 16 warps of
// three column groups each 186 us (cause not isolated: DESIGN.md section 8); profiles/lnbwd_tc_r02*.jsonl.)
// =================================================================================================
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 v) {
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

template <int NITER, int KS, int MODE, bool DOWN>
struct LnbTc {
  static constexpr int dim = NITER * 64, S = dim + 8, GW = NITER / 2, RP = KS * 8;
  static constexpr int S2 = dim + 16;                      // DOWN: the output projection's panel in tc_down's layout
  static constexpr int kSlots = GW * 2 * kTcThreads;       // one 16-byte (x, dres) or 8-byte (dy) slot per (column group, row, thread)
  static constexpr int kZ = kTcWarps * 16 * RP;            // DOWN: floats of one partial-product exchange buffer
  static constexpr size_t kSmem = ((size_t)RP * (DOWN ? S2 : S) + dim) * sizeof(float) + 2 * 16 * kTcWarps * sizeof(float2) + 2 * (size_t)kSlots * sizeof(float4) +
                                  (MODE == 0 ? (size_t)kSlots * sizeof(uint2) : 0) + (DOWN ? 2 * (size_t)kZ * sizeof(float) : 0);
};

// DOWN (MODE 0 without the additive term): the kernel also projects its OUTPUT rows, oz = dx @ ow^T (rank <= 8 KS) — the next consumer of
// the residual gradient in the GAViKO backward is the dgrad of Awakening_Prompt.proj_up (d(comb) = dG Wu, model/gaviko.py:187), which
// otherwise re-reads the 203 MB stream this kernel has just written.  The output registers are tc_down's A fragments; the per-warp
// partial products of a step are summed by the whole CTA after the NEXT step's barrier (no second barrier per step).
template <int NITER, int KS, int MODE, bool PGRAD, bool DOWN>
__global__ void __launch_bounds__(kTcThreads, 1) ln_bwd_tc_kernel(gvk_layernorm_bwd_params p) {
  using L = LnbTc<NITER, KS, MODE, DOWN>;
  constexpr int dim = L::dim, S = L::S, GW = L::GW, RP = L::RP, S2 = L::S2;
  constexpr bool RANK = !DOWN;                            // the rank-r product of MODE 0 / MODE 1
  static_assert(NITER % 2 == 0, "dim must be a multiple of 128");
  static_assert(!DOWN || (MODE == 0 && !PGRAD), "output projection: dense-dy form only");
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                                       // [RP][S] tf32 panel, two low column bits of every 16-column group swapped
  float* s_gamma = sW + RP * (DOWN ? S2 : S);             // [dim]
  float2* red = reinterpret_cast<float2*>(s_gamma + dim); // [2][16][kTcWarps]
  float4* s_x = reinterpret_cast<float4*>(red + 2 * 16 * kTcWarps);   // [GW][2][kTcThreads]
  float4* s_r = s_x + L::kSlots;
  uint2* s_y = reinterpret_cast<uint2*>(s_r + L::kSlots);  // MODE 0: raw bf16 x 4
  float* zbuf = reinterpret_cast<float*>(s_y + (MODE == 0 ? L::kSlots : 0));   // DOWN: [2][kTcWarps][16][RP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* rsrc = MODE == 0 ? p.az : p.dz;            // [M, r] latent operand of the rank-r product
  const int rld = MODE == 0 ? p.ld_az : p.ld_dz, rr = RANK ? (MODE == 0 ? p.ra : p.r) : 0;
  const int ntiles = (p.M + 15) / 16;
  const float inv_dim = 1.0f / dim;
  const int cw = warp * 16 * GW + 4 * t;                  // this lane's first column
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(p.dy);
  auto slot = [&](int kk, int row) { return (kk * 2 + row) * kTcThreads + tid; };
  auto row_a = [&](int tile) { return (size_t)min(tile * 16 + g, p.M - 1); };
  auto row_b = [&](int tile) { return (size_t)min(tile * 16 + g + 8, p.M - 1); };
  auto fetch_xy = [&](int tile, int kk) {
    const size_t cA = row_a(tile), cB = row_b(tile);
    cp_async16(s_x + slot(kk, 0), p.x + cA * p.ldx + cw + 16 * kk);
    cp_async16(s_x + slot(kk, 1), p.x + cB * p.ldx + cw + 16 * kk);
    if (MODE == 0) {
      cp_async8(s_y + slot(kk, 0), dyp + cA * p.ld_dy + cw + 16 * kk);
      cp_async8(s_y + slot(kk, 1), dyp + cB * p.ld_dy + cw + 16 * kk);
    }
  };
  auto fetch_r = [&](int tile, int kk) {
    cp_async16(s_r + slot(kk, 0), p.dres + row_a(tile) * p.ld_dres + cw + 16 * kk);
    cp_async16(s_r + slot(kk, 1), p.dres + row_b(tile) * p.ld_dres + cw + 16 * kk);
  };
  struct Lat { float v[KS][4]; float meanA, rstdA, meanB, rstdB; };
  auto load_lat = [&](int tile, Lat& T) {
    const size_t cA = row_a(tile), cB = row_b(tile);
#pragma unroll
    for (int s = 0; s < (RANK ? KS : 0); ++s) {
      const int k0 = 8 * s + t, k1 = k0 + 4;
      T.v[s][0] = k0 < rr ? rsrc[cA * rld + k0] : 0.f;
      T.v[s][1] = k0 < rr ? rsrc[cB * rld + k0] : 0.f;
      T.v[s][2] = k1 < rr ? rsrc[cA * rld + k1] : 0.f;
      T.v[s][3] = k1 < rr ? rsrc[cB * rld + k1] : 0.f;
    }
    T.meanA = p.mean[cA]; T.rstdA = p.rstd[cA]; T.meanB = p.mean[cB]; T.rstdB = p.rstd[cB];
  };
  // ---- prologue: the first step's streams start before the panel is staged
  Lat cur;
  if ((int)blockIdx.x < ntiles) {
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) fetch_xy(blockIdx.x, kk);
  }
  cp_async_commit();
  if ((int)blockIdx.x < ntiles && p.dres) {
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) fetch_r(blockIdx.x, kk);
  }
  cp_async_commit();
  if ((int)blockIdx.x < ntiles) load_lat(blockIdx.x, cur);
  if (DOWN) tc_stage_panel<RP, S2, false>(sW, p.ow, p.orank, dim, p.ow_sj, p.ow_sc, nullptr);
  else if (MODE == 0) tc_stage_panel<RP, S, true>(sW, p.aw, p.ra, dim, p.aw_sj, p.aw_sc, nullptr);
  else tc_stage_panel<RP, S, true>(sW, p.w, p.r, dim, p.w_sj, p.w_sc, nullptr);
  for (int c = tid; c < dim; c += kTcThreads) s_gamma[c] = p.gamma[c];
  __syncthreads();
  float4 dgm[PGRAD ? GW : 1], dbt[PGRAD ? GW : 1];
#pragma unroll
  for (int kk = 0; kk < (PGRAD ? GW : 1); ++kk) dgm[kk] = dbt[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
  // DOWN: oz rows of one finished step = sum of the 8 warps' partial products
  auto flush_oz = [&](int tile, const float* zb) {
    for (int idx = tid; idx < 16 * RP; idx += kTcThreads) {
      const int row = idx / RP, n = idx - row * RP, m = tile * 16 + row;
      float z = 0.f;
#pragma unroll
      for (int w = 0; w < kTcWarps; ++w) z += zb[w * 16 * RP + idx];
      if (n < p.orank && m < p.M) p.oz[(size_t)m * p.ld_oz + n] = z;
    }
  };
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int rA = tile * 16 + g, rB = rA + 8;
    const int next = tile + gridDim.x;
    const bool has_next = next < ntiles;                  // uniform over the CTA
    Lat nxt;
    if (has_next) load_lat(next, nxt);
    const float meanA = cur.meanA, rstdA = cur.rstdA, meanB = cur.meanB, rstdB = cur.rstdB;
    uint32_t a[KS][4];
#pragma unroll
    for (int s = 0; s < (RANK ? KS : 0); ++s)
#pragma unroll
      for (int e = 0; e < 4; ++e) a[s][e] = f2tf32(cur.v[s][e]);
    const float liveA = rA < p.M ? 1.f : 0.f, liveB = rB < p.M ? 1.f : 0.f;
    // rank-r product of one 16-column group on top of (accA, accB): lane's four consecutive columns of rows g and g+8
    auto rank_mma = [&](int kk, float4& accA, float4& accB) {
      float d0[4] = {accA.x, accA.y, accB.x, accB.y}, d1[4] = {accA.z, accA.w, accB.z, accB.w};
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const float2 b0 = *reinterpret_cast<const float2*>(sW + (8 * s + t) * S + 16 * (warp * GW + kk) + 2 * g);
        const float2 b1 = *reinterpret_cast<const float2*>(sW + (8 * s + t + 4) * S + 16 * (warp * GW + kk) + 2 * g);
        mma_tf32(d0, a[s][0], a[s][1], a[s][2], a[s][3], __float_as_uint(b0.x), __float_as_uint(b1.x));
        mma_tf32(d1, a[s][0], a[s][1], a[s][2], a[s][3], __float_as_uint(b0.y), __float_as_uint(b1.y));
      }
      accA = make_float4(d0[0], d0[1], d1[0], d1[1]);
      accB = make_float4(d0[2], d0[3], d1[2], d1[3]);
    };
    // ---- row statistics: m1 = mean(dy gamma), m2 = mean(dy gamma xhat).  Pending copy groups here: x / dy of this step, dres of this step.
    cp_async_wait<1>();
    float4 xa[GW], xb[GW];
    uint2 ya[MODE == 0 ? GW : 1], yb[MODE == 0 ? GW : 1];
    float s1A = 0.f, s2A = 0.f, s1B = 0.f, s2B = 0.f;
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) {
      xa[kk] = s_x[slot(kk, 0)];
      xb[kk] = s_x[slot(kk, 1)];
      if (MODE == 0) {
        ya[kk] = s_y[slot(kk, 0)];
        yb[kk] = s_y[slot(kk, 1)];
      }
      if (has_next) fetch_xy(next, kk);                   // the slots just read are free again
      const float4 gam = *reinterpret_cast<const float4*>(s_gamma + cw + 16 * kk);
      xa[kk] = make_float4((xa[kk].x - meanA) * rstdA, (xa[kk].y - meanA) * rstdA, (xa[kk].z - meanA) * rstdA, (xa[kk].w - meanA) * rstdA);
      xb[kk] = make_float4((xb[kk].x - meanB) * rstdB, (xb[kk].y - meanB) * rstdB, (xb[kk].z - meanB) * rstdB, (xb[kk].w - meanB) * rstdB);
      float4 dA, dB;
      if (MODE == 0) {
        dA = bf16x4_to_float4(ya[kk]);
        dB = bf16x4_to_float4(yb[kk]);
      } else {
        dA = dB = make_float4(0.f, 0.f, 0.f, 0.f);
        rank_mma(kk, dA, dB);
      }
      if (PGRAD) {
        dgm[kk].x += liveA * dA.x * xa[kk].x + liveB * dB.x * xb[kk].x;
        dgm[kk].y += liveA * dA.y * xa[kk].y + liveB * dB.y * xb[kk].y;
        dgm[kk].z += liveA * dA.z * xa[kk].z + liveB * dB.z * xb[kk].z;
        dgm[kk].w += liveA * dA.w * xa[kk].w + liveB * dB.w * xb[kk].w;
        dbt[kk].x += liveA * dA.x + liveB * dB.x;
        dbt[kk].y += liveA * dA.y + liveB * dB.y;
        dbt[kk].z += liveA * dA.z + liveB * dB.z;
        dbt[kk].w += liveA * dA.w + liveB * dB.w;
      }
      dA.x *= gam.x; dA.y *= gam.y; dA.z *= gam.z; dA.w *= gam.w;
      dB.x *= gam.x; dB.y *= gam.y; dB.z *= gam.z; dB.w *= gam.w;
      s1A += (dA.x + dA.y) + (dA.z + dA.w);
      s2A = fmaf(dA.x, xa[kk].x, fmaf(dA.y, xa[kk].y, fmaf(dA.z, xa[kk].z, fmaf(dA.w, xa[kk].w, s2A))));
      s1B += (dB.x + dB.y) + (dB.z + dB.w);
      s2B = fmaf(dB.x, xb[kk].x, fmaf(dB.y, xb[kk].y, fmaf(dB.z, xb[kk].z, fmaf(dB.w, xb[kk].w, s2B))));
    }
    cp_async_commit();                                    // x / dy of the next step
    s1A += __shfl_xor_sync(0xffffffffu, s1A, 1); s1A += __shfl_xor_sync(0xffffffffu, s1A, 2);
    s2A += __shfl_xor_sync(0xffffffffu, s2A, 1); s2A += __shfl_xor_sync(0xffffffffu, s2A, 2);
    s1B += __shfl_xor_sync(0xffffffffu, s1B, 1); s1B += __shfl_xor_sync(0xffffffffu, s1B, 2);
    s2B += __shfl_xor_sync(0xffffffffu, s2B, 1); s2B += __shfl_xor_sync(0xffffffffu, s2B, 2);
    float2* rbuf = red + buf * 16 * kTcWarps;
    if (t == 0) {
      rbuf[g * kTcWarps + warp] = make_float2(s1A, s2A);
      rbuf[(g + 8) * kTcWarps + warp] = make_float2(s1B, s2B);
    }
    __syncthreads();
    if (DOWN && tile != (int)blockIdx.x) flush_oz(tile - gridDim.x, zbuf + (buf ^ 1) * L::kZ);
    float m1A = 0.f, m2A = 0.f, m1B = 0.f, m2B = 0.f;
#pragma unroll
    for (int w = 0; w < kTcWarps; w += 2) {
      const float4 va = *reinterpret_cast<const float4*>(rbuf + g * kTcWarps + w);
      const float4 vb = *reinterpret_cast<const float4*>(rbuf + (g + 8) * kTcWarps + w);
      m1A += va.x + va.z; m2A += va.y + va.w;
      m1B += vb.x + vb.z; m2B += vb.y + vb.w;
    }
    m1A *= inv_dim; m2A *= inv_dim; m1B *= inv_dim; m2B *= inv_dim;
    // ---- dx = rstd (dy gamma - m1 - xhat m2) (+ az @ aw) (+ dres).  Pending copy groups: dres of this step, x / dy of the next.
    cp_async_wait<1>();
    float oacc[DOWN ? KS : 1][4];
#pragma unroll
    for (int j = 0; j < (DOWN ? KS : 1); ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) {
      const int col = cw + 16 * kk;
      float4 rcA = make_float4(0.f, 0.f, 0.f, 0.f), rcB = rcA;
      if (p.dres) {
        rcA = s_r[slot(kk, 0)];
        rcB = s_r[slot(kk, 1)];
        if (has_next) fetch_r(next, kk);
      }
      const float4 gam = *reinterpret_cast<const float4*>(s_gamma + col);
      float4 dA, dB;
      if (MODE == 0) {
        dA = bf16x4_to_float4(ya[kk]);
        dB = bf16x4_to_float4(yb[kk]);
      } else {
        dA = dB = make_float4(0.f, 0.f, 0.f, 0.f);
        rank_mma(kk, dA, dB);
      }
      float4 vA = make_float4(rstdA * (dA.x * gam.x - m1A - xa[kk].x * m2A), rstdA * (dA.y * gam.y - m1A - xa[kk].y * m2A),
                              rstdA * (dA.z * gam.z - m1A - xa[kk].z * m2A), rstdA * (dA.w * gam.w - m1A - xa[kk].w * m2A));
      float4 vB = make_float4(rstdB * (dB.x * gam.x - m1B - xb[kk].x * m2B), rstdB * (dB.y * gam.y - m1B - xb[kk].y * m2B),
                              rstdB * (dB.z * gam.z - m1B - xb[kk].z * m2B), rstdB * (dB.w * gam.w - m1B - xb[kk].w * m2B));
      if (MODE == 0 && RANK) rank_mma(kk, vA, vB);
      vA.x += rcA.x; vA.y += rcA.y; vA.z += rcA.z; vA.w += rcA.w;
      vB.x += rcB.x; vB.y += rcB.y; vB.z += rcB.z; vB.w += rcB.w;
      if (DOWN) {   // the finished output rows are tc_down's A fragments (k slot t <-> column 4t, t+4 <-> 4t+1; second MMA 4t+2 / 4t+3)
        const uint32_t ax = f2tf32(vA.x), ay = f2tf32(vA.y), az = f2tf32(vA.z), aw = f2tf32(vA.w);
        const uint32_t bx = f2tf32(vB.x), by = f2tf32(vB.y), bz = f2tf32(vB.z), bw = f2tf32(vB.w);
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(sW + (8 * j + g) * S2 + col);
          mma_tf32(oacc[j], ax, bx, ay, by, __float_as_uint(w.x), __float_as_uint(w.y));
          mma_tf32(oacc[j], az, bz, aw, bw, __float_as_uint(w.z), __float_as_uint(w.w));
        }
      }
      if (rA < p.M) {
        *reinterpret_cast<float4*>(p.dx + (size_t)rA * p.ld_dx + col) = vA;
        if (p.dx_lp) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(vA.x, vA.y), hi = __floats2bfloat162_rn(vA.z, vA.w);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.dx_lp) + (size_t)rA * p.ld_dx_lp + col) =
              make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
      }
      if (rB < p.M) {
        *reinterpret_cast<float4*>(p.dx + (size_t)rB * p.ld_dx + col) = vB;
        if (p.dx_lp) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(vB.x, vB.y), hi = __floats2bfloat162_rn(vB.z, vB.w);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.dx_lp) + (size_t)rB * p.ld_dx_lp + col) =
              make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
      }
    }
    cp_async_commit();                                    // dres of the next step
    if (DOWN) {
      float* zb = zbuf + buf * L::kZ;
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        *reinterpret_cast<float2*>(zb + (warp * 16 + g) * RP + 8 * j + 2 * t) = make_float2(oacc[j][0], oacc[j][1]);
        *reinterpret_cast<float2*>(zb + (warp * 16 + g + 8) * RP + 8 * j + 2 * t) = make_float2(oacc[j][2], oacc[j][3]);
      }
    }
    if (has_next) cur = nxt;
    if (DOWN && !has_next) {                              // last step of this CTA
      __syncthreads();
      flush_oz(tile, zbuf + buf * L::kZ);
    }
  }
  cp_async_wait<0>();
  if (PGRAD) {
    // lanes that share t hold the same columns for different rows: sum over g, then one atomic per column from the g == 0 lanes
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) {
      float v[8] = {dgm[kk].x, dgm[kk].y, dgm[kk].z, dgm[kk].w, dbt[kk].x, dbt[kk].y, dbt[kk].z, dbt[kk].w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 4);
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
      }
      if (g == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (p.dgamma) atomicAdd(p.dgamma + cw + 16 * kk + e, v[e]);
          if (p.dbeta) atomicAdd(p.dbeta + cw + 16 * kk + e, v[4 + e]);
        }
      }
    }
  }
}

template <int NITER, int KS, int MODE, bool PGRAD, bool DOWN = false>
static int ln_bwd_tc_launch(const gvk_layernorm_bwd_params* p, cudaStream_t stream) {
  constexpr size_t smem = LnbTc<NITER, KS, MODE, DOWN>::kSmem;
  static_assert(smem <= 227 * 1024, "layernorm_bwd_tc: shared memory");
  static const int attr = cudaFuncSetAttribute(ln_bwd_tc_kernel<NITER, KS, MODE, PGRAD, DOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr != cudaSuccess) return cuda_status((cudaError_t)attr, "layernorm_bwd_tc (smem attribute)");
  const int ntiles = (p->M + 15) / 16;
  ln_bwd_tc_kernel<NITER, KS, MODE, PGRAD, DOWN><<<std::max(1, std::min(ntiles, sm_count())), kTcThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("layernorm_bwd_tc");
  return GVK_OK;
}

// Forms the tensor-core kernel takes (everything else stays on the fp32 kernel): dense bf16 dy + additive rank term without parameter
// gradients (MODE 0), dense bf16 dy + a projection of the output rows (DOWN), or rank-r dy alone with both parameter gradients (MODE 1);
// dim 384 / 768; 16-byte aligned fp32 streams.
bool layernorm_bwd_tc_supported(const gvk_layernorm_bwd_params* p) {
  auto al = [](const void* q, uintptr_t m) { return (reinterpret_cast<uintptr_t>(q) & m) == 0; };
  if (p->dim != 384 && p->dim != 768) return false;      // dim 1024 (16 column groups per warp pair) does not fit the register file without spills
  if (p->ssf_scale) return false;
  if (!al(p->x, 15) || p->ldx % 4 != 0 || !al(p->dx, 15) || p->ld_dx % 4 != 0) return false;
  if (p->dres && (!al(p->dres, 15) || p->ld_dres % 4 != 0)) return false;
  if (p->dx_lp && (!al(p->dx_lp, 7) || p->ld_dx_lp % 4 != 0)) return false;
  const bool dense = p->dy && p->dy_dtype == GVK_BF16 && !p->dz && !p->dgamma && !p->dbeta && al(p->dy, 7) && p->ld_dy % 4 == 0;
  const bool mode0 = dense && p->az && !p->ow && p->ra <= 32;
  const bool down = dense && !p->az && p->ow && p->orank <= 24;   // a 32-row panel and its exchange buffers do not fit next to the stream slots
  const bool mode1 = !p->dy && p->dz && !p->az && !p->ow && p->dgamma && p->dbeta && p->r <= 32;
  return mode0 || mode1 || down;
}

int layernorm_bwd_tc(const gvk_layernorm_bwd_params* p, cudaStream_t stream) {
  const bool mode0 = p->dy != nullptr, down = p->ow != nullptr;
  const int r = down ? p->orank : (mode0 ? p->ra : p->r);
#define GVK_LNB_TC(NITER)                                                                                                              \
  if (down) return ln_bwd_tc_launch<NITER, 3, 0, false, true>(p, stream);                                                              \
  if (mode0) return r <= 24 ? ln_bwd_tc_launch<NITER, 3, 0, false>(p, stream) : ln_bwd_tc_launch<NITER, 4, 0, false>(p, stream);       \
  return r <= 24 ? ln_bwd_tc_launch<NITER, 3, 1, true>(p, stream) : ln_bwd_tc_launch<NITER, 4, 1, true>(p, stream);
  switch (p->dim / 64) {
    case 6: GVK_LNB_TC(6)
    case 12: GVK_LNB_TC(12)
    default:
      set_last_error("gvk_layernorm_bwd(tf32): dim %d", p->dim);
      return GVK_ERR_UNSUPPORTED;
  }
#undef GVK_LNB_TC
}

// =================================================================================================
// LayerNorm forward + rank-r down-projection of the SAME rows in one pass over x:
//   y = LN(x) gamma + beta (bf16),  mean / rstd saved;     z = act(x W^T + bias), pre = the pre-activation.
// In the GAViKO layer both FeedForward's LayerNorm (model/vision_transformer.py:30) and Awakening_Prompt.proj_down (model/gaviko.py:155-156)
// read the residual stream g_mid; as two kernels (layernorm_fwd 59 us + tc_down 54 us at M = 66 k) the 203 MB stream is read twice.
// Same CTA layout as ln_bwd_tc_kernel: a CTA owns 16 rows per step, warp w the columns [w dim/8, (w+1) dim/8); x of the next step
// travels by cp.async into thread-private slots.  The lane's x registers are tc_down's A fragments (k slot t <-> column 4t, t+4 <-> 4t+1,
// second MMA 4t+2 / 4t+3), so each warp accumulates a [16, r] partial over its columns; partial products and row sums (sum x, sum x^2)
// are exchanged through shared memory with one __syncthreads per step (double-buffered).
// =================================================================================================
template <int NITER, int NT>
struct LnFwdDown {
  static constexpr int dim = NITER * 64, S = dim + 16, GW = NITER / 2, RP = NT * 8;
  static constexpr int kSlots = GW * 2 * kTcThreads;
  static constexpr int kZ = kTcWarps * 16 * RP;            // floats of one partial-product exchange buffer
  static constexpr size_t kSmem = ((size_t)RP * S + 2 * dim + RP) * sizeof(float) + 2 * 16 * kTcWarps * sizeof(float2) + 2 * (size_t)kZ * sizeof(float) +
                                  (size_t)kSlots * sizeof(float4);
};

template <int NITER, int NT>
__global__ void __launch_bounds__(kTcThreads, 1) ln_fwd_down_tc_kernel(gvk_layernorm_fwd_down_params p) {
  using L = LnFwdDown<NITER, NT>;
  constexpr int dim = L::dim, S = L::S, GW = L::GW, RP = L::RP;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                                       // [RP][S] tf32 panel
  float* s_gamma = sW + RP * S;                           // [dim]
  float* s_beta = s_gamma + dim;                          // [dim]
  float* s_bias = s_beta + dim;                           // [RP]
  float2* red = reinterpret_cast<float2*>(s_bias + RP);   // [2][16][kTcWarps]  (sum x, sum x^2)
  float* zbuf = reinterpret_cast<float*>(red + 2 * 16 * kTcWarps);   // [2][kTcWarps][16][RP]
  float4* s_x = reinterpret_cast<float4*>(zbuf + 2 * L::kZ);          // [GW][2][kTcThreads]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (p.M + 15) / 16;
  const float inv_dim = 1.0f / dim;
  const int cw = warp * 16 * GW + 4 * t;
  auto slot = [&](int kk, int row) { return (kk * 2 + row) * kTcThreads + tid; };
  auto fetch_x = [&](int tile, int kk) {
    const size_t cA = (size_t)min(tile * 16 + g, p.M - 1), cB = (size_t)min(tile * 16 + g + 8, p.M - 1);
    cp_async16(s_x + slot(kk, 0), p.x + cA * p.ldx + cw + 16 * kk);
    cp_async16(s_x + slot(kk, 1), p.x + cB * p.ldx + cw + 16 * kk);
  };
  if ((int)blockIdx.x < ntiles) {
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) fetch_x(blockIdx.x, kk);
  }
  cp_async_commit();
  tc_stage_panel<RP, S, false>(sW, p.w, p.r, dim, p.w_sj, p.w_sc, nullptr);
  for (int c = tid; c < dim; c += kTcThreads) {
    s_gamma[c] = p.gamma[c];
    s_beta[c] = p.beta[c];
  }
  if (tid < RP) s_bias[tid] = (p.bias && tid < p.r) ? p.bias[tid] : 0.f;
  __syncthreads();
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y);
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int rA = tile * 16 + g, rB = rA + 8;
    const int next = tile + gridDim.x;
    const bool has_next = next < ntiles;
    cp_async_wait<0>();
    float4 xa[GW], xb[GW];
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    float sA = 0.f, qA = 0.f, sB = 0.f, qB = 0.f;
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) {
      xa[kk] = s_x[slot(kk, 0)];
      xb[kk] = s_x[slot(kk, 1)];
      if (has_next) fetch_x(next, kk);                    // the slots just read are free again
      sA += (xa[kk].x + xa[kk].y) + (xa[kk].z + xa[kk].w);
      qA = fmaf(xa[kk].x, xa[kk].x, fmaf(xa[kk].y, xa[kk].y, fmaf(xa[kk].z, xa[kk].z, fmaf(xa[kk].w, xa[kk].w, qA))));
      sB += (xb[kk].x + xb[kk].y) + (xb[kk].z + xb[kk].w);
      qB = fmaf(xb[kk].x, xb[kk].x, fmaf(xb[kk].y, xb[kk].y, fmaf(xb[kk].z, xb[kk].z, fmaf(xb[kk].w, xb[kk].w, qB))));
      const uint32_t ax = f2tf32(xa[kk].x), ay = f2tf32(xa[kk].y), az = f2tf32(xa[kk].z), aw = f2tf32(xa[kk].w);
      const uint32_t bx = f2tf32(xb[kk].x), by = f2tf32(xb[kk].y), bz = f2tf32(xb[kk].z), bw = f2tf32(xb[kk].w);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sW + (8 * j + g) * S + cw + 16 * kk);
        mma_tf32(acc[j], ax, bx, ay, by, __float_as_uint(w.x), __float_as_uint(w.y));
        mma_tf32(acc[j], az, bz, aw, bw, __float_as_uint(w.z), __float_as_uint(w.w));
      }
    }
    cp_async_commit();
    sA += __shfl_xor_sync(0xffffffffu, sA, 1); sA += __shfl_xor_sync(0xffffffffu, sA, 2);
    qA += __shfl_xor_sync(0xffffffffu, qA, 1); qA += __shfl_xor_sync(0xffffffffu, qA, 2);
    sB += __shfl_xor_sync(0xffffffffu, sB, 1); sB += __shfl_xor_sync(0xffffffffu, sB, 2);
    qB += __shfl_xor_sync(0xffffffffu, qB, 1); qB += __shfl_xor_sync(0xffffffffu, qB, 2);
    float2* rbuf = red + buf * 16 * kTcWarps;
    float* zb = zbuf + buf * L::kZ;
    if (t == 0) {
      rbuf[g * kTcWarps + warp] = make_float2(sA, qA);
      rbuf[(g + 8) * kTcWarps + warp] = make_float2(sB, qB);
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      *reinterpret_cast<float2*>(zb + (warp * 16 + g) * RP + 8 * j + 2 * t) = make_float2(acc[j][0], acc[j][1]);
      *reinterpret_cast<float2*>(zb + (warp * 16 + g + 8) * RP + 8 * j + 2 * t) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    float mA = 0.f, vA = 0.f, mB = 0.f, vB = 0.f;
#pragma unroll
    for (int w = 0; w < kTcWarps; w += 2) {
      const float4 a4 = *reinterpret_cast<const float4*>(rbuf + g * kTcWarps + w);
      const float4 b4 = *reinterpret_cast<const float4*>(rbuf + (g + 8) * kTcWarps + w);
      mA += a4.x + a4.z; vA += a4.y + a4.w;
      mB += b4.x + b4.z; vB += b4.y + b4.w;
    }
    mA *= inv_dim; mB *= inv_dim;
    const float rsA = rsqrtf(fmaxf(vA * inv_dim - mA * mA, 0.f) + p.eps), rsB = rsqrtf(fmaxf(vB * inv_dim - mB * mB, 0.f) + p.eps);
    if (warp == 0 && t == 0) {
      if (rA < p.M) { if (p.mean) p.mean[rA] = mA; if (p.rstd) p.rstd[rA] = rsA; }
      if (rB < p.M) { if (p.mean) p.mean[rB] = mB; if (p.rstd) p.rstd[rB] = rsB; }
    }
    // ---- the rank-r output of the step: 16 x RP sums over the 8 warps' partials
    for (int idx = tid; idx < 16 * RP; idx += kTcThreads) {
      const int row = idx / RP, n = idx - row * RP;
      float z = 0.f;
#pragma unroll
      for (int w = 0; w < kTcWarps; ++w) z += zb[w * 16 * RP + idx];
      const int m = tile * 16 + row;
      if (n < p.r && m < p.M) {
        z += s_bias[n];
        if (p.pre) p.pre[(size_t)m * p.ldz + n] = z;
        if (p.act == GVK_ROWACT_QUICKGELU) z = quick_gelu(z);
        else if (p.act == GVK_ROWACT_RELU) z = fmaxf(z, 0.f);
        p.z[(size_t)m * p.ldz + n] = z;
      }
    }
    // ---- y = (x - mean) rstd gamma + beta
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) {
      const int col = cw + 16 * kk;
      const float4 gam = *reinterpret_cast<const float4*>(s_gamma + col);
      const float4 bet = *reinterpret_cast<const float4*>(s_beta + col);
      if (rA < p.M) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf((xa[kk].x - mA) * rsA, gam.x, bet.x), fmaf((xa[kk].y - mA) * rsA, gam.y, bet.y));
        __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf((xa[kk].z - mA) * rsA, gam.z, bet.z), fmaf((xa[kk].w - mA) * rsA, gam.w, bet.w));
        *reinterpret_cast<uint2*>(yp + (size_t)rA * p.ldy + col) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
      if (rB < p.M) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf((xb[kk].x - mB) * rsB, gam.x, bet.x), fmaf((xb[kk].y - mB) * rsB, gam.y, bet.y));
        __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf((xb[kk].z - mB) * rsB, gam.z, bet.z), fmaf((xb[kk].w - mB) * rsB, gam.w, bet.w));
        *reinterpret_cast<uint2*>(yp + (size_t)rB * p.ldy + col) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
  }
  cp_async_wait<0>();
}

template <int NITER, int NT>
static int ln_fwd_down_launch(const gvk_layernorm_fwd_down_params* p, cudaStream_t stream) {
  constexpr size_t smem = LnFwdDown<NITER, NT>::kSmem;
  static_assert(smem <= 227 * 1024, "layernorm_fwd_down: shared memory");
  static const int attr = cudaFuncSetAttribute(ln_fwd_down_tc_kernel<NITER, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr != cudaSuccess) return cuda_status((cudaError_t)attr, "layernorm_fwd_down (smem attribute)");
  const int ntiles = (p->M + 15) / 16;
  ln_fwd_down_tc_kernel<NITER, NT><<<std::max(1, std::min(ntiles, sm_count())), kTcThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("layernorm_fwd_down");
  return GVK_OK;
}

int layernorm_fwd_down(const gvk_layernorm_fwd_down_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->x && p->gamma && p->beta && p->y && p->w && p->z && p->M > 0, "gvk_layernorm_fwd_down: null pointer");
  GVK_CHECK_ARG(p->dim == 384 || p->dim == 768, "gvk_layernorm_fwd_down: dim %d (384 or 768)", p->dim);
  GVK_CHECK_ARG(p->r >= 1 && p->r <= 32, "gvk_layernorm_fwd_down: r=%d (1..32)", p->r);
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && p->ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(p->y) & 7) == 0 && p->ldy % 4 == 0,
                "gvk_layernorm_fwd_down: x must be 16-byte and y 8-byte aligned, leading dimensions multiples of 4");
#define GVK_LFD(NITER) return p->r <= 24 ? ln_fwd_down_launch<NITER, 3>(p, stream) : ln_fwd_down_launch<NITER, 4>(p, stream);
  if (p->dim == 384) { GVK_LFD(6) }
  GVK_LFD(12)
#undef GVK_LFD
}


// =================================================================================================
// Rank-r up-projection followed by a rank-r down-projection of its OUTPUT rows, one pass over the [M, dim] stream:
//   out = res + drop_up(c W + b)                                   (gvk_rowproj_up)
//   z   = act(drop_dn(out) W2^T + b2),  pre = the pre-activation   (gvk_rowproj_down on the rows just written)
// GAViKO forward: LocalSelfAttention.proj_up + proj_drop + residual (model/gaviko.py:242-243, 301) followed by Awakening_Prompt.proj_down of
// the new local stream (model/gaviko.py:155-156 on ll); backward: d(loc) += d(ul) Wd followed by the dgrad of proj_up with the replayed
// proj_drop mask.  As two kernels the 197 MB stream is written and read again.  CTA layout of ln_bwd_tc_kernel: 16 rows per step, warp w
// owns columns [w dim/8, (w+1) dim/8); res of the next step travels by cp.async into thread-private slots; the output registers of the
// up-projection (tc_up's C fragments) are tc_down's A fragments; per-warp partial products meet in shared memory (one barrier per step).
// 8 streaming warps + 8 helper warps (masks of the next step, output sums of the finished one), see the kernel.
// =================================================================================================
template <int NITER>
struct UpDown {
  static constexpr int dim = NITER * 64, SU = dim + 8, SD = dim + 16, GW = NITER / 2, RP = 24, KS = 3;
  static constexpr int kSlots = GW * 2 * kTcThreads;
  static constexpr int kZ = kTcWarps * 16 * RP;
  static constexpr int kMaskHalves = (GW + 1) / 2;          // 16-bit words of one thread's mask nibbles (rows g / g+8 of GW column groups)
  static constexpr size_t kSmem = ((size_t)RP * SU + (size_t)RP * SD + dim + RP) * sizeof(float) + 2 * (size_t)kZ * sizeof(float) + (size_t)kSlots * sizeof(float4) +
                                  2 * (size_t)kMaskHalves * kTcThreads * sizeof(uint16_t);
};

// keep bits of the 4 consecutive elements starting at e (e % 4 == 0) under gvk_rowproj_up's mask rule: bit i = 1 iff element e + i is kept.
// u32_to_unit(r) >= p  <=>  (r >> 8) >= ceil(p 2^24)  (both sides are exact in fp32), i.e. r >= thr with thr = ceil(p 2^24) << 8.
__device__ __forceinline__ uint32_t tc_keep4(uint64_t seed, uint64_t e, uint32_t thr) {
  const uint64_t ctr = e >> 2;
  const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return (r.x >= thr ? 1u : 0u) | (r.y >= thr ? 2u : 0u) | (r.z >= thr ? 4u : 0u) | (r.w >= thr ? 8u : 0u);
}

// 16 warps: warps 0-7 stream and multiply as described above; warps 8-15 produce the dropout mask of the NEXT step (one Philox call per four
// elements — ~100 instructions that the 8 streaming warps could not hide at two warps per scheduler: 155 us per launch with the calls in line)
// as nibbles in shared memory, and sum the partial products of the finished step after the step's barrier.
template <int NITER>
__global__ void __launch_bounds__(2 * kTcThreads, 1) up_down_tc_kernel(gvk_rowproj_up_down_params p) {
  using L = UpDown<NITER>;
  constexpr int dim = L::dim, SU = L::SU, SD = L::SD, GW = L::GW, RP = L::RP, KS = L::KS, MH = L::kMaskHalves;
  extern __shared__ __align__(16) float smem[];
  float* sWu = smem;                                      // [RP][SU] up panel, two low column bits of every 16-column group swapped
  float* sWd = sWu + RP * SU;                             // [RP][SD] down panel (tc_down's layout)
  float* s_bu = sWd + RP * SD;                            // [dim]
  float* s_bd = s_bu + dim;                               // [RP]
  float* zbuf = s_bd + RP;                                // [2][kTcWarps][16][RP]
  float4* s_r = reinterpret_cast<float4*>(zbuf + 2 * L::kZ);   // [GW][2][kTcThreads]
  uint16_t* s_mk = reinterpret_cast<uint16_t*>(s_r + L::kSlots);   // [2][MH][kTcThreads] mask nibbles (byte kk: low = row g, high = row g + 8)
  const bool producer = threadIdx.x >= kTcThreads;
  const int tid = threadIdx.x & (kTcThreads - 1), warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;   // producer thread i mirrors streaming thread i
  // one mask comes from the producer warps: the up mask if there is one, else the down mask; a launch with both computes the second in line
  const bool mk_up = p.up_drop_p > 0.f, mk_dn = !mk_up && p.dn_drop_p > 0.f, dn_inline = mk_up && p.dn_drop_p > 0.f;
  const float mk_p = mk_up ? p.up_drop_p : p.dn_drop_p;
  const uint64_t mk_seed = salted_seed(mk_up ? p.up_seed : p.dn_seed, p.seed_salt), mk_off = mk_up ? p.up_offset : p.dn_offset;
  const uint32_t mk_thr = (uint32_t)ceilf(mk_p * 16777216.0f) << 8;
  const float keep_up = mk_up ? 1.0f / (1.0f - p.up_drop_p) : 1.f, keep_dn = p.dn_drop_p > 0.f ? 1.0f / (1.0f - p.dn_drop_p) : 1.f;
  const uint64_t seed_dn = salted_seed(p.dn_seed, p.seed_salt);
  const int ntiles = (p.M + 15) / 16;
  const int cw = warp * 16 * GW + 4 * t;
  auto slot = [&](int kk, int row) { return (kk * 2 + row) * kTcThreads + tid; };
  auto fetch_r = [&](int tile, int kk) {
    const size_t cA = (size_t)min(tile * 16 + g, p.M - 1), cB = (size_t)min(tile * 16 + g + 8, p.M - 1);
    cp_async16(s_r + slot(kk, 0), p.res + cA * p.ld_res + cw + 16 * kk);
    cp_async16(s_r + slot(kk, 1), p.res + cB * p.ld_res + cw + 16 * kk);
  };
  auto make_masks = [&](int tile, int b) {
    const size_t cA = (size_t)min(tile * 16 + g, p.M - 1), cB = (size_t)min(tile * 16 + g + 8, p.M - 1);
    uint32_t by[2 * MH];
#pragma unroll
    for (int kk = 0; kk < 2 * MH; ++kk) {
      by[kk] = 0u;
      if (kk < GW) by[kk] = tc_keep4(mk_seed, mk_off + cA * dim + cw + 16 * kk, mk_thr) | (tc_keep4(mk_seed, mk_off + cB * dim + cw + 16 * kk, mk_thr) << 4);
    }
#pragma unroll
    for (int h = 0; h < MH; ++h) s_mk[(b * MH + h) * kTcThreads + tid] = (uint16_t)(by[2 * h] | (by[2 * h + 1] << 8));
  };
  struct Lat { float v[KS][4]; };
  auto load_lat = [&](int tile, Lat& T) {
    const size_t cA = (size_t)min(tile * 16 + g, p.M - 1), cB = (size_t)min(tile * 16 + g + 8, p.M - 1);
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int k0 = 8 * s + t, k1 = k0 + 4;
      T.v[s][0] = k0 < p.r ? p.c[cA * p.ldc + k0] : 0.f;
      T.v[s][1] = k0 < p.r ? p.c[cB * p.ldc + k0] : 0.f;
      T.v[s][2] = k1 < p.r ? p.c[cA * p.ldc + k1] : 0.f;
      T.v[s][3] = k1 < p.r ? p.c[cB * p.ldc + k1] : 0.f;
    }
  };
  Lat cur;
  if (!producer) {
    if ((int)blockIdx.x < ntiles) {
      if (p.res) {
#pragma unroll
        for (int kk = 0; kk < GW; ++kk) fetch_r(blockIdx.x, kk);
      }
      load_lat(blockIdx.x, cur);
    }
    cp_async_commit();
    tc_stage_panel<RP, SU, true>(sWu, p.w, p.r, dim, p.w_sj, p.w_sc, nullptr);
    tc_stage_panel<RP, SD, false>(sWd, p.w2, p.r2, dim, p.w2_sj, p.w2_sc, nullptr);
    for (int c = tid; c < dim; c += kTcThreads) s_bu[c] = p.bias ? p.bias[c] : 0.f;
    if (tid < RP) s_bd[tid] = (p.bias2 && tid < p.r2) ? p.bias2[tid] : 0.f;
  } else if ((mk_up || mk_dn) && (int)blockIdx.x < ntiles) {
    make_masks(blockIdx.x, 0);
  }
  __syncthreads();
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    const bool has_next = next < ntiles;
    float* zb = zbuf + buf * L::kZ;
    if (producer) {
      if ((mk_up || mk_dn) && has_next) make_masks(next, buf ^ 1);
      __syncthreads();
      // ---- z rows of the finished step: sum of the 8 streaming warps' partial products
      for (int idx = tid; idx < 16 * RP; idx += kTcThreads) {
        const int row = idx / RP, n = idx - row * RP, m = tile * 16 + row;
        float z = 0.f;
#pragma unroll
        for (int w = 0; w < kTcWarps; ++w) z += zb[w * 16 * RP + idx];
        if (n < p.r2 && m < p.M) {
          z += s_bd[n];
          if (p.pre) p.pre[(size_t)m * p.ldz + n] = z;
          if (p.act == GVK_ROWACT_QUICKGELU) z = quick_gelu(z);
          else if (p.act == GVK_ROWACT_RELU) z = fmaxf(z, 0.f);
          p.z[(size_t)m * p.ldz + n] = z;
        }
      }
      continue;
    }
    const int rA = tile * 16 + g, rB = rA + 8;
    const size_t cA = (size_t)min(rA, p.M - 1), cB = (size_t)min(rB, p.M - 1);
    Lat nxt;
    if (has_next) load_lat(next, nxt);
    uint32_t a[KS][4];
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
      for (int e = 0; e < 4; ++e) a[s][e] = f2tf32(cur.v[s][e]);
    float oacc[KS][4];
#pragma unroll
    for (int j = 0; j < KS; ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f;
    uint32_t mk[MH];
#pragma unroll
    for (int h = 0; h < MH; ++h) mk[h] = (mk_up || mk_dn) ? s_mk[(buf * MH + h) * kTcThreads + tid] : 0xFFFFu;
    cp_async_wait<0>();
#pragma unroll
    for (int kk = 0; kk < GW; ++kk) {
      const int col = cw + 16 * kk;
      const uint32_t nib = mk[kk >> 1] >> (8 * (kk & 1));   // bits 0-3: row g, bits 4-7: row g + 8
      float4 rcA = make_float4(0.f, 0.f, 0.f, 0.f), rcB = rcA;
      if (p.res) {
        rcA = s_r[slot(kk, 0)];
        rcB = s_r[slot(kk, 1)];
        if (has_next) fetch_r(next, kk);                  // the slots just read are free again
      }
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const float2 b0 = *reinterpret_cast<const float2*>(sWu + (8 * s + t) * SU + 16 * (warp * GW + kk) + 2 * g);
        const float2 b1 = *reinterpret_cast<const float2*>(sWu + (8 * s + t + 4) * SU + 16 * (warp * GW + kk) + 2 * g);
        mma_tf32(d0, a[s][0], a[s][1], a[s][2], a[s][3], __float_as_uint(b0.x), __float_as_uint(b1.x));
        mma_tf32(d1, a[s][0], a[s][1], a[s][2], a[s][3], __float_as_uint(b0.y), __float_as_uint(b1.y));
      }
      const float4 bias = *reinterpret_cast<const float4*>(s_bu + col);
      float4 vA = make_float4(d0[0] + bias.x, d0[1] + bias.y, d1[0] + bias.z, d1[1] + bias.w);
      float4 vB = make_float4(d0[2] + bias.x, d0[3] + bias.y, d1[2] + bias.z, d1[3] + bias.w);
      if (mk_up) {
        vA.x *= (nib & 1u) ? keep_up : 0.f; vA.y *= (nib & 2u) ? keep_up : 0.f; vA.z *= (nib & 4u) ? keep_up : 0.f; vA.w *= (nib & 8u) ? keep_up : 0.f;
        vB.x *= (nib & 16u) ? keep_up : 0.f; vB.y *= (nib & 32u) ? keep_up : 0.f; vB.z *= (nib & 64u) ? keep_up : 0.f; vB.w *= (nib & 128u) ? keep_up : 0.f;
      }
      vA.x += rcA.x; vA.y += rcA.y; vA.z += rcA.z; vA.w += rcA.w;
      vB.x += rcB.x; vB.y += rcB.y; vB.z += rcB.z; vB.w += rcB.w;
      if (rA < p.M) *reinterpret_cast<float4*>(p.out + (size_t)rA * p.ld_out + col) = vA;
      if (rB < p.M) *reinterpret_cast<float4*>(p.out + (size_t)rB * p.ld_out + col) = vB;
      if (mk_dn) {
        vA.x *= (nib & 1u) ? keep_dn : 0.f; vA.y *= (nib & 2u) ? keep_dn : 0.f; vA.z *= (nib & 4u) ? keep_dn : 0.f; vA.w *= (nib & 8u) ? keep_dn : 0.f;
        vB.x *= (nib & 16u) ? keep_dn : 0.f; vB.y *= (nib & 32u) ? keep_dn : 0.f; vB.z *= (nib & 64u) ? keep_dn : 0.f; vB.w *= (nib & 128u) ? keep_dn : 0.f;
      } else if (dn_inline) {
        const float4 ma = tc_drop4(seed_dn, p.dn_offset + cA * dim + col, p.dn_drop_p, keep_dn);
        const float4 mb = tc_drop4(seed_dn, p.dn_offset + cB * dim + col, p.dn_drop_p, keep_dn);
        vA.x *= ma.x; vA.y *= ma.y; vA.z *= ma.z; vA.w *= ma.w;
        vB.x *= mb.x; vB.y *= mb.y; vB.z *= mb.z; vB.w *= mb.w;
      }
      // the output rows are tc_down's A fragments (k slot t <-> column 4t, t+4 <-> 4t+1; second MMA 4t+2 / 4t+3)
      const uint32_t ax = f2tf32(vA.x), ay = f2tf32(vA.y), az = f2tf32(vA.z), aw = f2tf32(vA.w);
      const uint32_t bx = f2tf32(vB.x), by = f2tf32(vB.y), bz = f2tf32(vB.z), bw = f2tf32(vB.w);
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sWd + (8 * j + g) * SD + col);
        mma_tf32(oacc[j], ax, bx, ay, by, __float_as_uint(w.x), __float_as_uint(w.y));
        mma_tf32(oacc[j], az, bz, aw, bw, __float_as_uint(w.z), __float_as_uint(w.w));
      }
    }
    cp_async_commit();
#pragma unroll
    for (int j = 0; j < KS; ++j) {
      *reinterpret_cast<float2*>(zb + (warp * 16 + g) * RP + 8 * j + 2 * t) = make_float2(oacc[j][0], oacc[j][1]);
      *reinterpret_cast<float2*>(zb + (warp * 16 + g + 8) * RP + 8 * j + 2 * t) = make_float2(oacc[j][2], oacc[j][3]);
    }
    __syncthreads();
    if (has_next) cur = nxt;
  }
  if (!producer) cp_async_wait<0>();
}

template <int NITER>
static int up_down_launch(const gvk_rowproj_up_down_params* p, cudaStream_t stream) {
  constexpr size_t smem = UpDown<NITER>::kSmem;
  static_assert(smem <= 227 * 1024, "rowproj_up_down: shared memory");
  static const int attr = cudaFuncSetAttribute(up_down_tc_kernel<NITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr != cudaSuccess) return cuda_status((cudaError_t)attr, "rowproj_up_down (smem attribute)");
  const int ntiles = (p->M + 15) / 16;
  up_down_tc_kernel<NITER><<<std::max(1, std::min(ntiles, sm_count())), 2 * kTcThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("rowproj_up_down");
  return GVK_OK;
}

int rowproj_up_down(const gvk_rowproj_up_down_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->c && p->w && p->out && p->w2 && p->z && p->M > 0, "gvk_rowproj_up_down: null pointer");
  GVK_CHECK_ARG(p->dim == 384 || p->dim == 768, "gvk_rowproj_up_down: dim %d (384 or 768)", p->dim);
  GVK_CHECK_ARG(p->r >= 1 && p->r <= 24 && p->r2 >= 1 && p->r2 <= 24, "gvk_rowproj_up_down: r=%d r2=%d (1..24)", p->r, p->r2);
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(p->out) & 15) == 0 && p->ld_out % 4 == 0 && (!p->res || ((reinterpret_cast<uintptr_t>(p->res) & 15) == 0 && p->ld_res % 4 == 0)),
                "gvk_rowproj_up_down: out / res must be 16-byte aligned with leading dimensions multiples of 4");
  GVK_CHECK_ARG(p->up_offset % 4 == 0 && p->dn_offset % 4 == 0, "gvk_rowproj_up_down: dropout offsets must be multiples of 4");
  return p->dim == 384 ? up_down_launch<6>(p, stream) : up_down_launch<12>(p, stream);
}

}  // namespace gvk
