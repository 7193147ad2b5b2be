// gvk_gaviko.cu — GAViKO-specific fused kernels: patch gather / token assembly, gated prompt fusion (Awakening_Prompt core)
// forward + backward, pooled head forward + backward, focal / cross-entropy loss.  See include/gvk.h for the contracts and
// the reference citations.  All fp32, warp-shuffle reductions, coalesced row accesses.
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {

// =================================================================================================
// patch gather / token rows
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) patch_gather_kernel(const float* __restrict__ img, int B, int C, int D, int H, int W, int fp, int ps, T* __restrict__ out) {
  const int W4 = W / 4;
  const size_t total = (size_t)B * C * D * H * W4;
  const int nd = D / fp, nh = H / ps, nw = W / ps;
  const int K = C * fp * ps * ps;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % W4);
    size_t rest = i / W4;
    const int y = (int)(rest % H); rest /= H;
    const int z = (int)(rest % D); rest /= D;
    const int c = (int)(rest % C);
    const int b = (int)(rest / C);
    const float4 v = *reinterpret_cast<const float4*>(img + i * 4);
    const int x = x4 * 4;
    const int wp = x / ps, kw = x - wp * ps;
    const int hp = y / ps, kh = y - hp * ps;
    const int dp = z / fp, kd = z - dp * fp;
    const size_t row = (size_t)b * nd * nh * nw + ((size_t)dp * nh + hp) * nw + wp;
    const int col = ((c * fp + kd) * ps + kh) * ps + kw;
    T* o = out + row * K + col;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(o) = v;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(o) = pk;
    }
  }
}

int patch_gather(const float* img, int B, int C, int D, int H, int W, int fp, int ps, void* patches, int out_dtype, cudaStream_t stream) {
  GVK_CHECK_ARG(img && patches && B > 0 && C > 0, "gvk_patch_gather: bad argument");
  GVK_CHECK_ARG(fp > 0 && ps > 0 && D % fp == 0 && H % ps == 0 && W % ps == 0 && ps % 4 == 0, "gvk_patch_gather: volume %dx%dx%d not divisible by patch %dx%dx%d (ps %% 4 == 0)", D,
                H, W, fp, ps, ps);
  const size_t total = (size_t)B * C * D * H * (W / 4);
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 32);
  if (out_dtype == GVK_F32)
    patch_gather_kernel<float><<<grid, 256, 0, stream>>>(img, B, C, D, H, W, fp, ps, reinterpret_cast<float*>(patches));
  else
    patch_gather_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(img, B, C, D, H, W, fp, ps, reinterpret_cast<__nv_bfloat16*>(patches));
  GVK_CHECK_LAUNCH("patch_gather");
  return GVK_OK;
}

__global__ void __launch_bounds__(256) fill_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, int R, int dim, float* __restrict__ out, int ld_out,
                                                          int out_batch_rows, int out_row_offset) {
  const int r = blockIdx.x, bb = blockIdx.y;
  float* o = out + ((size_t)bb * out_batch_rows + out_row_offset + r) * ld_out;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) o[c] = a[(size_t)r * dim + c] + (b ? b[(size_t)r * dim + c] : 0.f);
}
int fill_rows(const float* a, const float* b, int R, int dim, float* out, int ld_out, int out_batch_rows, int out_row_offset, int B, cudaStream_t stream) {
  GVK_CHECK_ARG(a && out && R > 0 && dim > 0 && B > 0, "gvk_fill_rows: bad argument");
  fill_rows_kernel<<<dim3(R, B), 256, 0, stream>>>(a, b, R, dim, out, ld_out, out_batch_rows, out_row_offset);
  GVK_CHECK_LAUNCH("fill_rows");
  return GVK_OK;
}

__global__ void __launch_bounds__(256) batch_rowsum_kernel(const float* __restrict__ x, int ldx, int batch_rows, int row_offset, int dim, int B, float* __restrict__ out,
                                                             int accumulate) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += x[((size_t)b * batch_rows + row_offset + r) * ldx + c];
    if (accumulate)
      out[(size_t)r * dim + c] += s;
    else
      out[(size_t)r * dim + c] = s;
  }
}
int batch_rowsum(const float* x, int ldx, int batch_rows, int row_offset, int R, int dim, int B, float* out, int accumulate, cudaStream_t stream) {
  GVK_CHECK_ARG(x && out && R > 0 && dim > 0 && B > 0, "gvk_batch_rowsum: bad argument");
  batch_rowsum_kernel<<<R, 256, 0, stream>>>(x, ldx, batch_rows, row_offset, dim, B, out, accumulate);
  GVK_CHECK_LAUNCH("batch_rowsum");
  return GVK_OK;
}

// =================================================================================================
// gated prompt fusion
// =================================================================================================
constexpr int kFusWarps = 8;
constexpr int kHid = 64;  // PromptRelevantEstimator hidden width (model/gaviko.py:25)

template <int R>
__device__ __forceinline__ float dotR(const float* __restrict__ row, const float (&q)[R]) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < R; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + c);
    s = fmaf(q[c], t.x, s); s = fmaf(q[c + 1], t.y, s); s = fmaf(q[c + 2], t.z, s); s = fmaf(q[c + 3], t.w, s);
  }
  return s;
}
template <int R>
__device__ __forceinline__ void axpyR(float a, const float* __restrict__ row, float (&acc)[R]) {
#pragma unroll
  for (int c = 0; c < R; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + c);
    acc[c] = fmaf(a, t.x, acc[c]); acc[c + 1] = fmaf(a, t.y, acc[c + 1]); acc[c + 2] = fmaf(a, t.z, acc[c + 2]); acc[c + 3] = fmaf(a, t.w, acc[c + 3]);
  }
}
template <int R>
__device__ __forceinline__ void load_vec_f4(const float* __restrict__ row, float (&v)[R]) {
#pragma unroll
  for (int c = 0; c < R; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + c);
    v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
  }
}
// lane j holds v[j] -> every lane gets the full vector
template <int R>
__device__ __forceinline__ void gatherR(float v_lane, float (&v)[R]) {
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = __shfl_sync(0xffffffffu, v_lane, j);
}
// lane j' (< R) returns sum_j W[j'*R + j] * v[j] + bias[j']
template <int R>
__device__ __forceinline__ float matvecR(const float* __restrict__ W, const float* __restrict__ bias, const float (&v)[R], int lane) {
  float acc = 0.f;
  if (lane < R) {
    acc = bias ? bias[lane] : 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) acc = fmaf(W[lane * R + j], v[j], acc);
  }
  return acc;
}
// LayerNorm over an R-vector held one element per lane (lanes >= R hold 0); returns the normalised (pre-affine) value, mean/rstd out
template <int R>
__device__ __forceinline__ float lnR(float x_lane, int lane, float& rstd) {
  const float mean = warp_sum(lane < R ? x_lane : 0.f) * (1.0f / R);
  const float d = lane < R ? x_lane - mean : 0.f;
  rstd = rsqrtf(warp_sum(d * d) * (1.0f / R) + 1e-5f);
  return d * rstd;
}

// softmax(q . tok) tok for one query held by a whole warp; q is pre-scaled.  Returns ctx[c] in lane c and the lse in every lane.
template <int R>
__device__ __forceinline__ float single_query_attn(const float* __restrict__ tok, int n, const float (&q)[R], int lane, float& lse) {
  float m = -INFINITY, l = 0.f;
  float acc[R];
#pragma unroll
  for (int c = 0; c < R; ++c) acc[c] = 0.f;
  for (int t = lane; t < n; t += 32) {
    const float* row = tok + (size_t)t * R;
    const float s = dotR<R>(row, q);
    if (s > m) {
      const float corr = __expf(m - s);
      l *= corr;
#pragma unroll
      for (int c = 0; c < R; ++c) acc[c] *= corr;
      m = s;
    }
    const float pr = __expf(s - m);
    l += pr;
    axpyR<R>(pr, row, acc);
  }
  const float M = warp_max(m);
  const float corr = (m == -INFINITY) ? 0.f : __expf(m - M);
  l = warp_sum(l * corr);
  const float inv_l = 1.0f / l;
  float out = 0.f;
#pragma unroll
  for (int c = 0; c < R; ++c) {
    const float v = warp_sum(acc[c] * corr) * inv_l;
    if (lane == c) out = v;
  }
  lse = M + __logf(l);
  return out;
}

// Gates from the cls latent (lane j holds cl[j]).  Returns imp for prompt p (all lanes) and gw (all lanes).
// Optionally exposes the intermediates needed by backward.
template <int R>
struct GateState {
  float cn_a, rstd_a;   // normalised (pre-affine) cls latent for the estimator LN, lane j
  float cn_g, rstd_g;   // same for the balancer LN (identical x-hat, kept separate for clarity)
  float hpre[2];        // hidden pre-activations k = lane, lane + 32
};
template <int R>
__device__ __forceinline__ void gates_hidden(const gvk_fusion_weights& w, float cl_lane, int lane, GateState<R>& st, float& gw) {
  const float xh = lnR<R>(cl_lane, lane, st.rstd_a);
  st.cn_a = xh;
  st.cn_g = xh;
  st.rstd_g = st.rstd_a;
  const float ya = lane < R ? xh * w.a_ln_w[lane] + w.a_ln_b[lane] : 0.f;
  const float yg = lane < R ? xh * w.g_ln_w[lane] + w.g_ln_b[lane] : 0.f;
  float ya_all[R];
  gatherR<R>(ya, ya_all);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int k = lane + 32 * u;
    float h = w.a_b1[k];
#pragma unroll
    for (int j = 0; j < R; ++j) h = fmaf(w.a_w1[k * R + j], ya_all[j], h);
    st.hpre[u] = h;
  }
  const float u = warp_sum(lane < R ? yg * w.g_w[lane] : 0.f) + w.g_b[0];
  gw = sigmoidf_(u);
}
template <int R>
__device__ __forceinline__ float gate_imp(const gvk_fusion_weights& w, const GateState<R>& st, int p, int lane) {
  float o = gelu_erf(st.hpre[0]) * w.a_w3[p * kHid + lane] + gelu_erf(st.hpre[1]) * w.a_w3[p * kHid + lane + 32];
  o = warp_sum(o) + w.a_b3[p];
  return sigmoidf_(o);
}

// Both key sets of a volume (xl[:, 2P+2:] and ll: 157 KB for T = 1033, N = 1000, r = 20) fit in shared memory next to the merge buffers,
// so the CTA copies them once with coalesced 16-byte loads and the per-key walk reads broadcast shared memory (~30 cycles) instead of L2
// (~600): with one warp-wide global load per key the walk was latency-bound at 67 - 84 us per call.  `staged` is 0 when they do not fit.
__device__ __forceinline__ void fus_stage_keys(float* dst, const float* __restrict__ src, int nfloats) {
#pragma unroll 4
  for (int i = threadIdx.x * 4; i < nfloats; i += blockDim.x * 4) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(src + i);
}
static bool fus_can_stage(const float* xl, const float* ll, int T, int N, int P, int r, size_t base_bytes, size_t* total) {
  const size_t keys = ((size_t)(T - 2 * P - 2) + N) * r * sizeof(float);
  const bool aligned = r % 4 == 0 && (((size_t)T * r) % 4 == 0) && ((reinterpret_cast<uintptr_t>(xl) | reinterpret_cast<uintptr_t>(ll)) & 15) == 0;
  const bool ok = aligned && base_bytes + keys <= 227 * 1024;
  *total = base_bytes + (ok ? keys : 0);
  return ok;
}

// Forward: one CTA per volume.  A lane owns one prompt (query, running max / sum and context accumulator in registers), a warp walks a
// slice of the keys of one of the two cross-attentions (warps [0, W/2): global keys xl[:, 2P+2:], warps [W/2, W): local keys ll) reading
// each key row once as a warp-wide broadcast; the W/2 partial softmax states of a side are merged through shared memory.  Warp 0 computes
// the gates (they depend on the cls latent only) before it joins the key walk.
// (History: one warp per (b, prompt) walking all ~2000 keys with a shuffle reduction per context element was pure latency, 54 us.)
constexpr int kFusFwdWarps = 16;
template <int R>
struct FusFwdSmem {
  static constexpr int kPart = kFusFwdWarps * (R + 2) * 32;   // part[warp][R acc | m | l][prompt]
  static constexpr int kW = 2 * (R * R + R);                  // wq_g | bq_g | wq_l | bq_l
  static constexpr int kCtx = 2 * 32 * (R + 1);               // ctx[side][prompt][c]
  static constexpr int kGate = 32 + 4 + kHid;                 // imp[prompt of the group], gw, gelu(hidden) of the estimator
  static constexpr size_t kBytes = (size_t)(kPart + kW + kCtx + kGate) * sizeof(float);
};
template <int R>
__global__ void __launch_bounds__(kFusFwdWarps * 32) fusion_fwd_kernel(gvk_fusion_fwd_params p, int staged) {
  extern __shared__ __align__(16) float fus_smem[];
  float* part = fus_smem;
  float* s_w = part + FusFwdSmem<R>::kPart;
  float* s_ctx = s_w + FusFwdSmem<R>::kW;
  float* s_imp = s_ctx + FusFwdSmem<R>::kCtx;
  float* s_keys = s_imp + FusFwdSmem<R>::kGate;
  constexpr int HW = kFusFwdWarps / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int side = warp / HW, wl = warp % HW;
  const int b = blockIdx.x;
  float* xl_b = p.xl + (size_t)b * p.T * R;
  const float scale = rsqrtf((float)R);
  for (int i = threadIdx.x; i < R * R + R; i += blockDim.x) {
    s_w[i] = i < R * R ? p.w.wq_g[i] : p.w.bq_g[i - R * R];
    s_w[R * R + R + i] = i < R * R ? p.w.wq_l[i] : p.w.bq_l[i - R * R];
  }
  const int n_g = p.T - 2 * p.P - 2;
  const int n = side == 0 ? n_g : p.N;
  const float* tokbase = side == 0 ? xl_b + (size_t)(2 * p.P + 2) * R : p.ll + (size_t)b * p.N * R;
  if (staged) {
    fus_stage_keys(s_keys, xl_b + (size_t)(2 * p.P + 2) * R, n_g * R);
    fus_stage_keys(s_keys + n_g * R, p.ll + (size_t)b * p.N * R, p.N * R);
    tokbase = side == 0 ? s_keys : s_keys + n_g * R;
  }
  const int per = (n + HW - 1) / HW;
  const int t0 = wl * per, t1 = min(n, t0 + per);
  GateState<R> gst;
  float gw = 0.f;
  if (warp == 0) {
    const float cl_lane = lane < R ? xl_b[(size_t)p.P * R + lane] : 0.f;
    gates_hidden<R>(p.w, cl_lane, lane, gst, gw);
    if (lane == 0) {
      s_imp[32] = gw;
      p.s.gw[b] = gw;
    }
  }
  __syncthreads();
  const float* wq = s_w + side * (R * R + R);
  for (int pg = 0; pg < p.P; pg += 32) {
    const int pp = pg + lane;
    const bool valid = pp < p.P;
    const size_t o = ((size_t)b * p.P + (valid ? pp : 0)) * R;
    if (warp == 0) {                                  // importance gate of this lane's prompt: sigmoid(W3[pp, :] . gelu(hidden) + b3[pp])
      float* s_hact = s_imp + 36;                     // gelu(hidden), written once, read by every group of 32 prompts
      if (pg == 0) {
        s_hact[lane] = gelu_erf(gst.hpre[0]);
        s_hact[lane + 32] = gelu_erf(gst.hpre[1]);
      }
      __syncwarp();
      if (valid) {
        const float* w3 = p.w.a_w3 + (size_t)pp * kHid;
        float o0 = p.w.a_b3[pp], o1 = 0.f;
#pragma unroll
        for (int k = 0; k < kHid; k += 8) {
          const float4 wa = *reinterpret_cast<const float4*>(w3 + k), wb = *reinterpret_cast<const float4*>(w3 + k + 4);
          o0 = fmaf(wa.x, s_hact[k], fmaf(wa.y, s_hact[k + 1], fmaf(wa.z, s_hact[k + 2], fmaf(wa.w, s_hact[k + 3], o0))));
          o1 = fmaf(wb.x, s_hact[k + 4], fmaf(wb.y, s_hact[k + 5], fmaf(wb.z, s_hact[k + 6], fmaf(wb.w, s_hact[k + 7], o1))));
        }
        const float imp = sigmoidf_(o0 + o1);
        s_imp[lane] = imp;
        p.s.imp[(size_t)b * p.P + pp] = imp;
      }
    }
    float q[R];
    {
      float pl[R];
      load_vec_f4<R>(xl_b + (size_t)(valid ? pp : 0) * R, pl);
#pragma unroll
      for (int c = 0; c < R; ++c) {
        float v = wq[R * R + c];
#pragma unroll
        for (int j = 0; j < R; ++j) v = fmaf(wq[c * R + j], pl[j], v);
        q[c] = v;
      }
      if (valid && wl == 0) {
        float* qdst = (side == 0 ? p.s.qg : p.s.ql) + o;
#pragma unroll
        for (int c = 0; c < R; ++c) {
          qdst[c] = q[c];
          if (side == 0) p.s.pl[o + c] = pl[c];
        }
      }
#pragma unroll
      for (int c = 0; c < R; ++c) q[c] *= scale;
    }
    float m = -INFINITY, l = 0.f;
    float acc[R];
#pragma unroll
    for (int c = 0; c < R; ++c) acc[c] = 0.f;
#pragma unroll(R <= 20 ? 2 : 1)
    for (int t = t0; t < t1; ++t) {
      float tok[R];
      load_vec_f4<R>(tokbase + (size_t)t * R, tok);      // the same address in every lane: one broadcast transaction
      float sc = 0.f;
#pragma unroll
      for (int c = 0; c < R; ++c) sc = fmaf(q[c], tok[c], sc);
      const float mn = fmaxf(m, sc);
      const float corr = __expf(m - mn), pr = __expf(sc - mn);   // exp(-inf) = 0 on the first key
      m = mn;
      l = fmaf(l, corr, pr);
#pragma unroll
      for (int c = 0; c < R; ++c) acc[c] = fmaf(acc[c], corr, pr * tok[c]);
    }
    float* mine = part + (size_t)warp * (R + 2) * 32 + lane;
#pragma unroll
    for (int c = 0; c < R; ++c) mine[c * 32] = acc[c];
    mine[R * 32] = m;
    mine[(R + 1) * 32] = l;
    __syncthreads();   // also orders every warp's read of the prompt rows before they are overwritten below
    if (wl == 0) {
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < HW; ++w) M = fmaxf(M, part[((side * HW + w) * (R + 2) + R) * 32 + lane]);
      float L = 0.f;
      float ctx[R];
#pragma unroll
      for (int c = 0; c < R; ++c) ctx[c] = 0.f;
#pragma unroll
      for (int w = 0; w < HW; ++w) {
        const float* pw = part + (size_t)(side * HW + w) * (R + 2) * 32 + lane;
        const float mw = pw[R * 32];
        const float f = mw == -INFINITY ? 0.f : __expf(mw - M);
        L = fmaf(pw[(R + 1) * 32], f, L);
#pragma unroll
        for (int c = 0; c < R; ++c) ctx[c] = fmaf(pw[c * 32], f, ctx[c]);
      }
      const float inv_l = 1.0f / L;
      if (valid) {
        float* cdst = (side == 0 ? p.s.ctx_g : p.s.ctx_l) + o;
#pragma unroll
        for (int c = 0; c < R; ++c) {
          ctx[c] *= inv_l;
          cdst[c] = ctx[c];
          s_ctx[(side * 32 + lane) * (R + 1) + c] = ctx[c];
        }
        (side == 0 ? p.s.lse_g : p.s.lse_l)[(size_t)b * p.P + pp] = M + __logf(L);
      }
    }
    __syncthreads();
    const float gwv = s_imp[32];
    for (int idx = threadIdx.x; idx < 32 * R; idx += blockDim.x) {   // combined_latent row p = (gw ctx_g + (1 - gw) ctx_l) imp_p
      const int k = idx / R, c = idx - k * R;
      if (pg + k < p.P) xl_b[(size_t)(pg + k) * R + c] = (gwv * s_ctx[k * (R + 1) + c] + (1.f - gwv) * s_ctx[(32 + k) * (R + 1) + c]) * s_imp[k];
    }
    __syncthreads();   // shared buffers are reused by the next group of 32 prompts
  }
}

// ---- backward A: one warp per (b, prompt) ------------------------------------------------------
// ws layout per (b,p): [dctx_g (R) | dctx_l (R) | delta_g | delta_l | dimp | dgw]
template <int R>
__device__ __forceinline__ float single_query_attn_dq(const float* __restrict__ tok, int t_begin, int t_end, const float (&q)[R], const float (&dctx)[R], float lse,
                                                      float delta, int lane) {
  float dq[R];
#pragma unroll
  for (int c = 0; c < R; ++c) dq[c] = 0.f;
  for (int t = t_begin + lane; t < t_end; t += 32) {
    const float* row = tok + (size_t)t * R;
    const float a = __expf(dotR<R>(row, q) - lse);
    const float ds = a * (dotR<R>(row, dctx) - delta);
    axpyR<R>(ds, row, dq);
  }
  float out = 0.f;
#pragma unroll
  for (int c = 0; c < R; ++c) {
    const float v = warp_sum(dq[c]);
    if (lane == c) out = v;
  }
  return out;  // lane c: sum_j ds_j tok_j[c]   (caller multiplies by the softmax scale)
}

// One CTA per volume.  A lane owns one prompt (its pre-scaled query, d ctx, lse and delta live in registers), a warp walks a slice of the
// keys of one of the two cross-attentions (warps [0, W/2): global keys xl[:, 2P+2:], warps [W/2, W): local keys ll) reading each key row
// once as a warp-wide broadcast, so dQ needs no shuffles at all; the W/2 partial dQ of a side are summed through shared memory and the
// query-projection gradients leave the CTA as one atomicAdd per weight (B x 840 atomics per call instead of B x P x 840).
// (History: one warp per (b, prompt) was pure latency, 185 us; one CTA per (b, prompt) with the keys split over 8 warps was 108 us for
// 41 MFMA at B = 32 — 2 CTAs / SM by registers, 7 waves, 1.7 M contended atomics.)
constexpr int kFusBwdWarps = 16;
template <int R>
struct FusBwdSmem {
  static constexpr int kPart = kFusBwdWarps * R * 32;     // part[warp][c][prompt]
  static constexpr int kDq = 2 * 32 * (R + 1);            // dq[side][prompt][c], scaled
  static constexpr int kPl = 32 * (R + 1);                // pl[prompt][j]
  static constexpr size_t kBytes = (size_t)(kPart + kDq + kPl) * sizeof(float);
};
template <int R>
__global__ void __launch_bounds__(kFusBwdWarps * 32) fusion_bwd_prompts_kernel(gvk_fusion_bwd_params p, int staged) {
  extern __shared__ __align__(16) float fus_smem[];
  float* part = fus_smem;
  float* s_dq = part + FusBwdSmem<R>::kPart;
  float* s_pl = s_dq + FusBwdSmem<R>::kDq;
  float* s_keys = s_pl + FusBwdSmem<R>::kPl;
  constexpr int HW = kFusBwdWarps / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int side = warp / HW, wl = warp % HW;
  const int b = blockIdx.x;
  const float scale = rsqrtf((float)R);
  const float gw = p.s.gw[b];
  float* dxl_b = p.dxl + (size_t)b * p.T * R;
  const int n_g = p.T - 2 * p.P - 2;
  const int n = side == 0 ? n_g : p.N;
  const float* tokbase = side == 0 ? p.xl + ((size_t)b * p.T + 2 * p.P + 2) * R : p.ll + (size_t)b * p.N * R;
  if (staged) {
    fus_stage_keys(s_keys, p.xl + ((size_t)b * p.T + 2 * p.P + 2) * R, n_g * R);
    fus_stage_keys(s_keys + n_g * R, p.ll + (size_t)b * p.N * R, p.N * R);
    tokbase = side == 0 ? s_keys : s_keys + n_g * R;
    __syncthreads();
  }
  const int per = (n + HW - 1) / HW;
  const int t0 = wl * per, t1 = min(n, t0 + per);
  for (int pg = 0; pg < p.P; pg += 32) {
    const int pp = pg + lane;
    const bool valid = pp < p.P;
    const size_t o = ((size_t)b * p.P + (valid ? pp : 0)) * R;
    float q[R], dctx[R];
    float lse = INFINITY, delta = 0.f;
    {
      float cg[R], cl[R];
      load_vec_f4<R>(p.s.ctx_g + o, cg);
      load_vec_f4<R>(p.s.ctx_l + o, cl);
      load_vec_f4<R>(dxl_b + (size_t)(valid ? pp : 0) * R, dctx);   // d enh
      const float imp = p.s.imp[(size_t)b * p.P + (valid ? pp : 0)];
      float d_imp = 0.f, d_gw = 0.f;
#pragma unroll
      for (int c = 0; c < R; ++c) {
        d_imp = fmaf(dctx[c], gw * cg[c] + (1.f - gw) * cl[c], d_imp);
        const float d_fused = dctx[c] * imp;
        d_gw = fmaf(d_fused, cg[c] - cl[c], d_gw);
        dctx[c] = (side == 0 ? gw : 1.f - gw) * d_fused;
        delta = fmaf(dctx[c], side == 0 ? cg[c] : cl[c], delta);
      }
      load_vec_f4<R>((side == 0 ? p.s.qg : p.s.ql) + o, q);
      if (valid) {
        lse = (side == 0 ? p.s.lse_g : p.s.lse_l)[(size_t)b * p.P + pp];
        if (wl == 0) {                                                   // hand-off to the token / gate kernels
          float* ws = p.ws + ((size_t)b * p.P + pp) * (2 * R + 4);
#pragma unroll
          for (int c = 0; c < R; ++c) ws[side * R + c] = dctx[c];
          ws[2 * R + side] = delta;
          if (side == 0) {
            ws[2 * R + 2] = d_imp;
            ws[2 * R + 3] = d_gw;
            load_vec_f4<R>(p.s.pl + o, cg);
#pragma unroll
            for (int c = 0; c < R; ++c) s_pl[lane * (R + 1) + c] = cg[c];
          }
        }
      } else if (wl == 0 && side == 0) {
#pragma unroll
        for (int c = 0; c < R; ++c) s_pl[lane * (R + 1) + c] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < R; ++c) q[c] = valid ? q[c] * scale : 0.f;
    }
    float dq[R];
#pragma unroll
    for (int c = 0; c < R; ++c) dq[c] = 0.f;
#pragma unroll(R <= 20 ? 4 : 1)
    for (int t = t0; t < t1; ++t) {
      float tok[R];
      load_vec_f4<R>(tokbase + (size_t)t * R, tok);                      // the same address in every lane: one broadcast transaction
      float sc = 0.f, da = 0.f;
#pragma unroll
      for (int c = 0; c < R; ++c) {
        sc = fmaf(q[c], tok[c], sc);
        da = fmaf(dctx[c], tok[c], da);
      }
      const float ds = __expf(sc - lse) * (da - delta);
#pragma unroll
      for (int c = 0; c < R; ++c) dq[c] = fmaf(ds, tok[c], dq[c]);
    }
#pragma unroll
    for (int c = 0; c < R; ++c) part[(warp * R + c) * 32 + lane] = dq[c];
    __syncthreads();   // also orders every warp's read of the d enh rows before they are overwritten below
    if (wl == 0) {
#pragma unroll
      for (int c = 0; c < R; ++c) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < HW; ++w) v += part[((side * HW + w) * R + c) * 32 + lane];
        s_dq[(side * 32 + lane) * (R + 1) + c] = valid ? v * scale : 0.f;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * R * (R + 1); idx += blockDim.x) {   // d Wq[c][j] and (j == R) d bq[c] of both sides
      const int sd = idx / (R * (R + 1)), c = (idx / (R + 1)) % R, j = idx % (R + 1);
      float v = 0.f;
      for (int k = 0; k < 32; ++k) v = fmaf(s_dq[(sd * 32 + k) * (R + 1) + c], j < R ? s_pl[k * (R + 1) + j] : 1.f, v);
      float* dst = j < R ? (sd == 0 ? p.g.wq_g : p.g.wq_l) + c * R + j : (sd == 0 ? p.g.bq_g : p.g.bq_l) + c;
      atomicAdd(dst, v);
    }
    for (int idx = threadIdx.x; idx < 32 * R; idx += blockDim.x) {            // dL/d(prompt latent) = Wq_g^T dq_g + Wq_l^T dq_l
      const int k = idx / R, j = idx % R;
      if (pg + k < p.P) {
        float v = 0.f;
#pragma unroll
        for (int c = 0; c < R; ++c) v = fmaf(p.w.wq_g[c * R + j], s_dq[k * (R + 1) + c], fmaf(p.w.wq_l[c * R + j], s_dq[(32 + k) * (R + 1) + c], v));
        dxl_b[(size_t)(pg + k) * R + j] = v;
      }
    }
    __syncthreads();   // shared buffers are reused by the next group of 32 prompts
  }
}

// ---- backward B: one thread per key token ------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) fusion_bwd_tokens_kernel(gvk_fusion_bwd_params p) {
  extern __shared__ float sm[];  // per prompt: q (R, pre-scaled), dctx (R), lse, delta
  const int b = blockIdx.y;
  const int which = blockIdx.z;  // 0 global, 1 local
  float* sq = sm;
  float* sd = sm + p.P * R;
  float* sl = sd + p.P * R;
  float* sdel = sl + p.P;
  const float scale = rsqrtf((float)R);
  for (int i = threadIdx.x; i < p.P * R; i += blockDim.x) {
    const int pp = i / R, c = i - pp * R;
    const size_t o = ((size_t)b * p.P + pp) * R + c;
    sq[i] = (which == 0 ? p.s.qg[o] : p.s.ql[o]) * scale;
    sd[i] = p.ws[((size_t)b * p.P + pp) * (2 * R + 4) + which * R + c];
  }
  for (int pp = threadIdx.x; pp < p.P; pp += blockDim.x) {
    sl[pp] = which == 0 ? p.s.lse_g[(size_t)b * p.P + pp] : p.s.lse_l[(size_t)b * p.P + pp];
    sdel[pp] = p.ws[((size_t)b * p.P + pp) * (2 * R + 4) + 2 * R + which];
  }
  __syncthreads();
  const int n = which == 0 ? p.T - 2 * p.P - 2 : p.N;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float* row = which == 0 ? p.xl + ((size_t)b * p.T + 2 * p.P + 2 + t) * R : p.ll + ((size_t)b * p.N + t) * R;
  float tok[R], dt[R];
#pragma unroll
  for (int c = 0; c < R; c += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    tok[c] = v.x; tok[c + 1] = v.y; tok[c + 2] = v.z; tok[c + 3] = v.w;
  }
#pragma unroll
  for (int c = 0; c < R; ++c) dt[c] = 0.f;
  for (int pp = 0; pp < p.P; ++pp) {
    const float* q = sq + pp * R;
    const float* dc = sd + pp * R;
    float s = 0.f, da = 0.f;
#pragma unroll
    for (int c = 0; c < R; ++c) {
      s = fmaf(q[c], tok[c], s);
      da = fmaf(dc[c], tok[c], da);
    }
    const float a = __expf(s - sl[pp]);
    const float ds = a * (da - sdel[pp]);
#pragma unroll
    for (int c = 0; c < R; ++c) dt[c] += a * dc[c] + ds * q[c];  // q already carries the softmax scale
  }
  if (which == 0) {
    float* d = p.dxl + ((size_t)b * p.T + 2 * p.P + 2 + t) * R;
#pragma unroll
    for (int c = 0; c < R; ++c) d[c] += dt[c];
  } else {
    float* d = p.dll + ((size_t)b * p.N + t) * R;
#pragma unroll
    for (int c = 0; c < R; ++c) d[c] = dt[c];
  }
}

// ---- backward C: gates, one warp per volume ----------------------------------------------------
template <int R>
__global__ void __launch_bounds__(32) fusion_bwd_gates_kernel(gvk_fusion_bwd_params p) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const float cl_lane = lane < R ? p.xl[((size_t)b * p.T + p.P) * R + lane] : 0.f;
  GateState<R> st;
  float gw;
  gates_hidden<R>(p.w, cl_lane, lane, st, gw);
  const float hact[2] = {gelu_erf(st.hpre[0]), gelu_erf(st.hpre[1])};
  float d_hact[2] = {0.f, 0.f};
  float d_gw = 0.f;
  // per-prompt terms are loaded 32 prompts at a time, one per lane, before the dependent chain of atomics starts
  for (int pg = 0; pg < p.P; pg += 32) {
    const int pl = pg + lane;
    float d_o_lane = 0.f;
    if (pl < p.P) {
      const float* ws = p.ws + ((size_t)b * p.P + pl) * (2 * R + 4);
      const float imp = p.s.imp[(size_t)b * p.P + pl];
      d_o_lane = ws[2 * R + 2] * imp * (1.f - imp);
      d_gw += ws[2 * R + 3];
    }
    const int cnt = min(32, p.P - pg);
    float w3[2][32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int pp = pg + min(k, cnt - 1);
      w3[0][k] = p.w.a_w3[pp * kHid + lane];
      w3[1][k] = p.w.a_w3[pp * kHid + lane + 32];
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < cnt) {
        const float d_o = __shfl_sync(0xffffffffu, d_o_lane, k);
        atomicAdd(p.g.a_w3 + (pg + k) * kHid + lane, d_o * hact[0]);
        atomicAdd(p.g.a_w3 + (pg + k) * kHid + lane + 32, d_o * hact[1]);
        d_hact[0] = fmaf(d_o, w3[0][k], d_hact[0]);
        d_hact[1] = fmaf(d_o, w3[1][k], d_hact[1]);
      }
    }
    if (pl < p.P) atomicAdd(p.g.a_b3 + pl, d_o_lane);
  }
  d_gw = warp_sum(d_gw);
  // estimator: hidden -> LN_a(cl)
  const float ya = lane < R ? st.cn_a * p.w.a_ln_w[lane] + p.w.a_ln_b[lane] : 0.f;
  float ya_all[R];
  gatherR<R>(ya, ya_all);
  float d_ya[R];
#pragma unroll
  for (int j = 0; j < R; ++j) d_ya[j] = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int k = lane + 32 * u;
    const float d_hpre = d_hact[u] * gelu_erf_grad(st.hpre[u]);
    atomicAdd(p.g.a_b1 + k, d_hpre);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      atomicAdd(p.g.a_w1 + k * R + j, d_hpre * ya_all[j]);
      d_ya[j] = fmaf(d_hpre, p.w.a_w1[k * R + j], d_ya[j]);
    }
  }
  float d_ya_lane = 0.f;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const float v = warp_sum(d_ya[j]);
    if (lane == j) d_ya_lane = v;
  }
  // balancer
  const float d_u = d_gw * gw * (1.f - gw);
  const float yg = lane < R ? st.cn_g * p.w.g_ln_w[lane] + p.w.g_ln_b[lane] : 0.f;
  float d_cl = 0.f;
  if (lane < R) {
    atomicAdd(p.g.g_w + lane, d_u * yg);
    atomicAdd(p.g.a_ln_w + lane, d_ya_lane * st.cn_a);
    atomicAdd(p.g.a_ln_b + lane, d_ya_lane);
  }
  if (lane == 0) atomicAdd(p.g.g_b, d_u);
  const float d_yg_lane = lane < R ? d_u * p.w.g_w[lane] : 0.f;
  if (lane < R) {
    atomicAdd(p.g.g_ln_w + lane, d_yg_lane * st.cn_g);
    atomicAdd(p.g.g_ln_b + lane, d_yg_lane);
  }
  // LayerNorm backward of both branches (same x-hat / rstd)
  {
    const float g1 = lane < R ? d_ya_lane * p.w.a_ln_w[lane] + d_yg_lane * p.w.g_ln_w[lane] : 0.f;
    const float m1 = warp_sum(g1) * (1.0f / R);
    const float m2 = warp_sum(g1 * st.cn_a) * (1.0f / R);
    d_cl = st.rstd_a * (g1 - m1 - st.cn_a * m2);
  }
  if (lane < R) p.dxl[((size_t)b * p.T + p.P) * R + lane] += d_cl;
}

template <int R>
static int fusion_fwd_launch(const gvk_fusion_fwd_params* p, cudaStream_t stream) {
  static const int attr = cudaFuncSetAttribute(fusion_fwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (attr != cudaSuccess) return cuda_status((cudaError_t)attr, "prompt_fusion_fwd (smem attribute)");
  size_t smem;
  const int staged = fus_can_stage(p->xl, p->ll, p->T, p->N, p->P, R, FusFwdSmem<R>::kBytes, &smem);
  fusion_fwd_kernel<R><<<p->B, kFusFwdWarps * 32, smem, stream>>>(*p, staged);
  GVK_CHECK_LAUNCH("prompt_fusion_fwd");
  return GVK_OK;
}
template <int R>
static int fusion_bwd_launch(const gvk_fusion_bwd_params* p, cudaStream_t stream) {
  static const int attr = cudaFuncSetAttribute(fusion_bwd_prompts_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (attr != cudaSuccess) return cuda_status((cudaError_t)attr, "prompt_fusion_bwd (smem attribute)");
  size_t smem;
  const int staged = fus_can_stage(p->xl, p->ll, p->T, p->N, p->P, R, FusBwdSmem<R>::kBytes, &smem);
  fusion_bwd_prompts_kernel<R><<<p->B, kFusBwdWarps * 32, smem, stream>>>(*p, staged);
  GVK_CHECK_LAUNCH("prompt_fusion_bwd_prompts");
  const int nmax = std::max(p->N, p->T - 2 * p->P - 2);
  const size_t smem_tok = ((size_t)2 * p->P * R + 2 * p->P) * sizeof(float);
  fusion_bwd_tokens_kernel<R><<<dim3((nmax + 255) / 256, p->B, 2), 256, smem_tok, stream>>>(*p);
  GVK_CHECK_LAUNCH("prompt_fusion_bwd_tokens");
  fusion_bwd_gates_kernel<R><<<p->B, 32, 0, stream>>>(*p);
  GVK_CHECK_LAUNCH("prompt_fusion_bwd_gates");
  return GVK_OK;
}

static int check_fusion(int B, int T, int N, int P, int r, const char* who) {
  GVK_CHECK_ARG(B > 0 && P > 0 && N > 0 && T > 2 * P + 2, "%s: bad shape B=%d T=%d N=%d P=%d (needs T > 2P+2)", who, B, T, N, P);
  GVK_CHECK_ARG(r == 16 || r == 20 || r == 32, "%s: latent dim %d unsupported (16, 20, 32)", who, r);
  GVK_CHECK_ARG((size_t)P * (2 * r + 2) * sizeof(float) <= 48 * 1024, "%s: too many prompts (%d) for the token-gradient kernel", who, P);
  return GVK_OK;
}

int prompt_fusion_fwd(const gvk_fusion_fwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->xl && p->ll, "gvk_prompt_fusion_fwd: null pointer");
  int st = check_fusion(p->B, p->T, p->N, p->P, p->r, "gvk_prompt_fusion_fwd");
  if (st != GVK_OK) return st;
  switch (p->r) {
    case 16: return fusion_fwd_launch<16>(p, stream);
    case 20: return fusion_fwd_launch<20>(p, stream);
    default: return fusion_fwd_launch<32>(p, stream);
  }
}
int prompt_fusion_bwd(const gvk_fusion_bwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->xl && p->ll && p->dxl && p->dll && p->ws, "gvk_prompt_fusion_bwd: null pointer");
  int st = check_fusion(p->B, p->T, p->N, p->P, p->r, "gvk_prompt_fusion_bwd");
  if (st != GVK_OK) return st;
  switch (p->r) {
    case 16: return fusion_bwd_launch<16>(p, stream);
    case 20: return fusion_bwd_launch<20>(p, stream);
    default: return fusion_bwd_launch<32>(p, stream);
  }
}

__global__ void __launch_bounds__(256) quickgelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = dy[i] * quick_gelu_grad(pre[i]);
}
int quickgelu_bwd(const float* dy, const float* pre, float* y, size_t n, cudaStream_t stream) {
  GVK_CHECK_ARG(dy && pre && y && n > 0, "gvk_quickgelu_bwd: bad argument");
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 8);
  quickgelu_bwd_kernel<<<grid, 256, 0, stream>>>(dy, pre, y, n);
  GVK_CHECK_LAUNCH("quickgelu_bwd");
  return GVK_OK;
}

// =================================================================================================
// head
// =================================================================================================
constexpr int kHeadWarps = 8;

__global__ void __launch_bounds__(kHeadWarps * 32) head_fwd_kernel(gvk_head_fwd_params p) {
  extern __shared__ float sm[];  // [kHeadWarps][dim] partial sums, then pooled [dim]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dim = p.dim;
  float* part = sm + warp * dim;
  for (int c = lane; c < dim; c += 32) part[c] = 0.f;
  const float inv_dim = 1.0f / dim;
  for (int r = warp; r < p.pool_count; r += kHeadWarps) {
    const float* x = p.x + ((size_t)b * p.T + p.pool_start + r) * p.ldx;
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s += x[c];
    const float mean = warp_sum(s) * inv_dim;
    float v = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float d = x[c] - mean;
      v += d * d;
    }
    const float rstd = rsqrtf(warp_sum(v) * inv_dim + p.eps);
    for (int c = lane; c < dim; c += 32) {
      float y = (x[c] - mean) * rstd * p.gamma[c] + p.beta[c];
      if (p.ssf_scale) y = y * p.ssf_scale[c] + p.ssf_shift[c];
      part[c] += y;
    }
  }
  __syncthreads();
  float* pooled = sm + kHeadWarps * dim;
  const float inv_cnt = 1.0f / p.pool_count;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kHeadWarps; ++w) s += sm[w * dim + c];
    s *= inv_cnt;
    pooled[c] = s;
    p.pooled[(size_t)b * dim + c] = s;
  }
  __syncthreads();
  for (int k = warp; k < p.num_classes; k += kHeadWarps) {
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s = fmaf(pooled[c], p.wh[(size_t)k * dim + c], s);
    s = warp_sum(s);
    if (lane == 0) p.logits[(size_t)b * p.num_classes + k] = s + p.bh[k];
  }
}

int head_fwd(const gvk_head_fwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->x && p->gamma && p->beta && p->wh && p->bh && p->pooled && p->logits, "gvk_head_fwd: null pointer");
  GVK_CHECK_ARG(p->B > 0 && p->dim > 0 && p->dim <= 2048 && p->pool_count > 0 && p->pool_start >= 0 && p->pool_start + p->pool_count <= p->T && p->num_classes > 0,
                "gvk_head_fwd: bad shape");
  const size_t smem = (size_t)(kHeadWarps + 1) * p->dim * sizeof(float);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    int st = cuda_status(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "head_fwd smem");
    if (st != GVK_OK) return st;
    configured = smem;
  }
  head_fwd_kernel<<<p->B, kHeadWarps * 32, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("head_fwd");
  return GVK_OK;
}

__global__ void __launch_bounds__(kHeadWarps * 32) head_bwd_rows_kernel(gvk_head_bwd_params bp) {
  extern __shared__ float sm[];  // dpooled [dim] (already divided by pool_count and multiplied by ssf_scale)
  const gvk_head_fwd_params& p = bp.f;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dim = p.dim;
  const float inv_cnt = 1.0f / p.pool_count;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < p.num_classes; ++k) s = fmaf(bp.dlogits[(size_t)b * p.num_classes + k], p.wh[(size_t)k * dim + c], s);
    s *= inv_cnt;  // d(LN-out [* ssf]) per pooled row
    if (bp.dssf_shift) atomicAdd(bp.dssf_shift + c, s * p.pool_count);
    sm[c] = s;
  }
  __syncthreads();
  const float inv_dim = 1.0f / dim;
  for (int r = warp; r < p.pool_count; r += kHeadWarps) {
    const size_t row = (size_t)b * p.T + p.pool_start + r;
    const float* x = p.x + row * p.ldx;
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s += x[c];
    const float mean = warp_sum(s) * inv_dim;
    float v = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float d = x[c] - mean;
      v += d * d;
    }
    const float rstd = rsqrtf(warp_sum(v) * inv_dim + p.eps);
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float xh = (x[c] - mean) * rstd;
      float dy = sm[c];
      if (p.ssf_scale) {
        if (bp.dssf_scale) atomicAdd(bp.dssf_scale + c, dy * (xh * p.gamma[c] + p.beta[c]));
        dy *= p.ssf_scale[c];
      }
      if (bp.dgamma) atomicAdd(bp.dgamma + c, dy * xh);
      if (bp.dbeta) atomicAdd(bp.dbeta + c, dy);
      const float g = dy * p.gamma[c];
      s1 += g;
      s2 += g * xh;
    }
    const float m1 = warp_sum(s1) * inv_dim, m2 = warp_sum(s2) * inv_dim;
    for (int c = lane; c < dim; c += 32) {
      const float xh = (x[c] - mean) * rstd;
      float dy = sm[c];
      if (p.ssf_scale) dy *= p.ssf_scale[c];
      const float dx = rstd * (dy * p.gamma[c] - m1 - xh * m2);
      bp.dx[row * bp.ld_dx + c] = dx;
      if (bp.dx_lp) reinterpret_cast<__nv_bfloat16*>(bp.dx_lp)[row * bp.ld_dx_lp + c] = __float2bfloat16_rn(dx);
    }
  }
}

__global__ void __launch_bounds__(256) head_bwd_w_kernel(gvk_head_bwd_params bp) {
  const gvk_head_fwd_params& p = bp.f;
  const int k = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < p.dim) {
    float s = 0.f;
    for (int b = 0; b < p.B; ++b) s = fmaf(bp.dlogits[(size_t)b * p.num_classes + k], p.pooled[(size_t)b * p.dim + c], s);
    if (bp.accumulate_w) bp.dwh[(size_t)k * p.dim + c] += s; else bp.dwh[(size_t)k * p.dim + c] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < p.B; ++b) s += bp.dlogits[(size_t)b * p.num_classes + k];
    if (bp.accumulate_w) bp.dbh[k] += s; else bp.dbh[k] = s;
  }
}

int head_bwd(const gvk_head_bwd_params* bp, cudaStream_t stream) {
  GVK_CHECK_ARG(bp && bp->dlogits && bp->f.x && bp->f.pooled && bp->f.wh, "gvk_head_bwd: null pointer");
  const gvk_head_fwd_params& p = bp->f;
  if (bp->dx) {
    head_bwd_rows_kernel<<<p.B, kHeadWarps * 32, p.dim * sizeof(float), stream>>>(*bp);
    GVK_CHECK_LAUNCH("head_bwd_rows");
  }
  if (bp->dwh && bp->dbh) {
    head_bwd_w_kernel<<<dim3((p.dim + 255) / 256, p.num_classes), 256, 0, stream>>>(*bp);
    GVK_CHECK_LAUNCH("head_bwd_w");
  }
  return GVK_OK;
}

// =================================================================================================
// losses
// =================================================================================================
constexpr int kMaxClasses = 32;

__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int B, int C, int kind, float gamma, float eps,
                                                     long long ignore_index, float* __restrict__ loss, float* __restrict__ dlogits) {
  __shared__ float s_loss[256];
  __shared__ int s_cnt[256];
  float my_loss = 0.f;
  int my_cnt = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) my_cnt += (target[b] != ignore_index) ? 1 : 0;
  s_cnt[threadIdx.x] = my_cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
    __syncthreads();
  }
  const float inv_cnt = 1.0f / (float)s_cnt[0];
  const float hi = 1.0f - eps;  // == 1.0f in fp32 for eps = 1e-16, exactly as in the reference
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* z = logits + (size_t)b * C;
    float* dz = dlogits ? dlogits + (size_t)b * C : nullptr;
    const long long y = target[b];
    if (y == ignore_index) {
      if (dz)
        for (int k = 0; k < C; ++k) dz[k] = 0.f;
      continue;
    }
    if (y < 0 || y >= C) {   // out-of-range label: the reference's one_hot / CrossEntropyLoss raise; a device kernel cannot, so poison the result
      my_loss = __int_as_float(0x7fc00000);
      if (dz)
        for (int k = 0; k < C; ++k) dz[k] = __int_as_float(0x7fc00000);
      continue;
    }
    if (kind == 1) {  // cross entropy
      float m = -INFINITY;
      for (int k = 0; k < C; ++k) m = fmaxf(m, z[k]);
      float s = 0.f;
      for (int k = 0; k < C; ++k) s += expf(z[k] - m);
      const float lse = m + logf(s);
      my_loss += lse - z[y];
      if (dz)
        for (int k = 0; k < C; ++k) dz[k] = (expf(z[k] - lse) - (k == y ? 1.f : 0.f)) * inv_cnt;
      continue;
    }
    float p1[kMaxClasses], p2[kMaxClasses];
    bool pass1[kMaxClasses], pass2[kMaxClasses];
    float m = -INFINITY;
    for (int k = 0; k < C; ++k) {
      pass1[k] = (z[k] >= eps) && (z[k] <= hi);
      p1[k] = fminf(fmaxf(z[k], eps), hi);
      m = fmaxf(m, p1[k]);
    }
    float s = 0.f;
    for (int k = 0; k < C; ++k) {
      p1[k] = expf(p1[k] - m);
      s += p1[k];
    }
    m = -INFINITY;
    for (int k = 0; k < C; ++k) {
      p1[k] /= s;
      pass2[k] = (p1[k] >= eps) && (p1[k] <= hi);
      p2[k] = fminf(fmaxf(p1[k], eps), hi);
      m = fmaxf(m, p2[k]);
    }
    s = 0.f;
    for (int k = 0; k < C; ++k) {
      p2[k] = expf(p2[k] - m);
      s += p2[k];
    }
    for (int k = 0; k < C; ++k) p2[k] /= s;
    const float pt = p2[y];
    const float nll = -logf(eps + pt);
    const float foc = powf(1.f - pt, gamma);
    my_loss += foc * nll;
    if (dz) {
      const float dpt = (-gamma * powf(1.f - pt, gamma - 1.f) * nll - foc / (eps + pt)) * inv_cnt;
      // softmax #2 backward, clamp #2, softmax #1 backward, clamp #1
      float dc2[kMaxClasses];
      float dot = 0.f;
      for (int k = 0; k < C; ++k) {
        dc2[k] = p2[k] * ((k == y ? 1.f : 0.f) - pt) * dpt;
        if (!pass2[k]) dc2[k] = 0.f;
        dot += p1[k] * dc2[k];
      }
      for (int k = 0; k < C; ++k) dz[k] = pass1[k] ? p1[k] * (dc2[k] - dot) : 0.f;
    }
  }
  s_loss[threadIdx.x] = my_loss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_loss[threadIdx.x] += s_loss[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = s_loss[0] * inv_cnt;
}

int loss_fwd_bwd(const float* logits, const long long* target, int B, int C, int kind, float gamma, float eps, long long ignore_index, float* loss, float* dlogits,
                 cudaStream_t stream) {
  GVK_CHECK_ARG(logits && target && loss && B > 0, "gvk_loss_fwd_bwd: bad argument");
  GVK_CHECK_ARG(C >= 2 && C <= kMaxClasses, "gvk_loss_fwd_bwd: C=%d must be in [2,%d]", C, kMaxClasses);
  GVK_CHECK_ARG(kind == 0 || kind == 1, "gvk_loss_fwd_bwd: unknown loss kind %d", kind);
  loss_kernel<<<1, 256, 0, stream>>>(logits, target, B, C, kind, gamma, eps, ignore_index, loss, dlogits);
  GVK_CHECK_LAUNCH("loss_fwd_bwd");
  return GVK_OK;
}

}  // namespace gvk
