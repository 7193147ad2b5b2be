// gvk_attn_simt.cu — exact-fp32 softmax attention on CUDA cores (see include/gvk.h: gvk_attn_simt_fwd / _bwd).
// Window-sparse (GAViKO LocalSelfAttention, model/gaviko.py:212-244) and dense (fp32-mode MHSA, model/vision_transformer.py:65-70).
// Forward and dQ: one warp per query, lanes stride over the query's keys with a per-lane online softmax merged at the end.
// dK/dV: one warp per key, lanes stride over the queries that see this key (the transposed window is again a box), so no atomics.
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {

constexpr int kAttnWarps = 4;

struct Box {  // inclusive key/query box of one token in the 3-D token grid
  int d0, h0, w0, nd, nh, nw;
  __device__ __forceinline__ int count() const { return nd * nh * nw; }
  __device__ __forceinline__ int token(int t, int gh, int gw) const {
    const int tw = t % nw;
    const int th = (t / nw) % nh;
    const int td = t / (nw * nh);
    return ((d0 + td) * gh + (h0 + th)) * gw + (w0 + tw);
  }
};
// keys of query i (lo = k/2, hi = k-1-k/2) or queries of key j (lo = k-1-k/2, hi = k/2)
__device__ __forceinline__ Box make_box(const gvk_attn_fwd_params& p, int tok, bool transposed) {
  const int w = tok % p.grid_w, h = (tok / p.grid_w) % p.grid_h, d = tok / (p.grid_w * p.grid_h);
  const int lod = transposed ? p.win_d - 1 - p.win_d / 2 : p.win_d / 2, hid = p.win_d - 1 - lod;
  const int loh = transposed ? p.win_h - 1 - p.win_h / 2 : p.win_h / 2, hih = p.win_h - 1 - loh;
  const int low = transposed ? p.win_w - 1 - p.win_w / 2 : p.win_w / 2, hiw = p.win_w - 1 - low;
  Box b;
  b.d0 = max(0, d - lod);
  b.h0 = max(0, h - loh);
  b.w0 = max(0, w - low);
  b.nd = min(p.grid_d - 1, d + hid) - b.d0 + 1;
  b.nh = min(p.grid_h - 1, h + hih) - b.h0 + 1;
  b.nw = min(p.grid_w - 1, w + hiw) - b.w0 + 1;
  return b;
}

template <int D>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c);
    v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
  }
}
template <int D>
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float (&v)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const uint2 t = *reinterpret_cast<const uint2*>(p + c);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[c] = a.x; v[c + 1] = a.y; v[c + 2] = b.x; v[c + 3] = b.y;
  }
}
template <int D>
__device__ __forceinline__ float dot_row(const float* p, const float (&q)[D]) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c);
    s = fmaf(q[c], t.x, s); s = fmaf(q[c + 1], t.y, s); s = fmaf(q[c + 2], t.z, s); s = fmaf(q[c + 3], t.w, s);
  }
  return s;
}
template <int D>
__device__ __forceinline__ float dot_row(const __nv_bfloat16* p, const float (&q)[D]) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const uint2 t = *reinterpret_cast<const uint2*>(p + c);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    s = fmaf(q[c], a.x, s); s = fmaf(q[c + 1], a.y, s); s = fmaf(q[c + 2], b.x, s); s = fmaf(q[c + 3], b.y, s);
  }
  return s;
}
template <int D>
__device__ __forceinline__ void axpy_row(float a, const float* p, float (&acc)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c);
    acc[c] = fmaf(a, t.x, acc[c]); acc[c + 1] = fmaf(a, t.y, acc[c + 1]); acc[c + 2] = fmaf(a, t.z, acc[c + 2]); acc[c + 3] = fmaf(a, t.w, acc[c + 3]);
  }
}
template <int D>
__device__ __forceinline__ void axpy_row(float a, const __nv_bfloat16* p, float (&acc)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const uint2 t = *reinterpret_cast<const uint2*>(p + c);
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    acc[c] = fmaf(a, x.x, acc[c]); acc[c + 1] = fmaf(a, x.y, acc[c + 1]); acc[c + 2] = fmaf(a, y.x, acc[c + 2]); acc[c + 3] = fmaf(a, y.y, acc[c + 3]);
  }
}
// Halving exchange (see gvk_rowops.cu): sums N per-lane partials across the warp with N - N/32 shuffles instead of 5 N.
// N = 32: lane L ends with the total of v[L] in v[0];  N = 64: totals of v[2L], v[2L+1] in v[0], v[1].
template <int N>
__device__ __forceinline__ void reduce_half_step(float* v, int lane, int o) {
  const bool upper = (lane & o) != 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float keep = upper ? v[k + N] : v[k];
    const float send = upper ? v[k] : v[k + N];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
  }
}
__device__ __forceinline__ void warp_reduce_scatter32(float (&v)[32], int lane) {
  reduce_half_step<16>(v, lane, 16); reduce_half_step<8>(v, lane, 8); reduce_half_step<4>(v, lane, 4); reduce_half_step<2>(v, lane, 2); reduce_half_step<1>(v, lane, 1);
}
__device__ __forceinline__ void warp_reduce_scatter64(float (&v)[64], int lane) {
  reduce_half_step<32>(v, lane, 16); reduce_half_step<16>(v, lane, 8); reduce_half_step<8>(v, lane, 4); reduce_half_step<4>(v, lane, 2); reduce_half_step<2>(v, lane, 1);
}
__device__ __forceinline__ void store_elem(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_elem(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float prob_drop(const gvk_attn_fwd_params& p, uint64_t seed, uint64_t elem, float inv_keep) {
  const uint64_t e = p.offset + elem;
  const uint64_t ctr = e >> 2;
  const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t sel = (uint32_t)(e & 3);
  const uint32_t v = sel == 0 ? r.x : sel == 1 ? r.y : sel == 2 ? r.z : r.w;
  return u32_to_unit(v) >= p.drop_p ? inv_keep : 0.f;
}

// ------------------------------------------------------------------------------------------------
template <int D, typename T>
__global__ void __launch_bounds__(kAttnWarps * 32) attn_fwd_kernel(gvk_attn_fwd_params p) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * kAttnWarps + warp;
  if (idx >= (long long)p.B * p.H * p.T) return;
  const int i = (int)(idx % p.T);
  const int bh = (int)(idx / p.T);
  const int h = bh % p.H, b = bh / p.H;
  const T* base = reinterpret_cast<const T*>(p.qkv) + (size_t)b * p.T * p.ld + h * D;
  float q[D];
  load_vec<D>(base + (size_t)i * p.ld + p.q_off, q);
#pragma unroll
  for (int c = 0; c < D; ++c) q[c] *= p.scale;
  const bool windowed = p.win_d > 0;
  Box box;
  int count = p.T;
  if (windowed) {
    box = make_box(p, i, false);
    count = box.count();
  }
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  float m = -INFINITY, l = 0.f;
  float acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.f;
  for (int t = lane; t < count; t += 32) {
    const int j = windowed ? box.token(t, p.grid_h, p.grid_w) : t;
    const T* kr = base + (size_t)j * p.ld + p.k_off;
    const float s = dot_row<D>(kr, q);
    if (s > m) {
      const float corr = __expf(m - s);  // exp(-inf) = 0 on first key
      l *= corr;
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] *= corr;
      m = s;
    }
    float pr = __expf(s - m);
    l += pr;
    if (p.drop_p > 0.f) pr *= prob_drop(p, seed_eff, ((uint64_t)bh * p.T + i) * p.T + j, inv_keep);
    axpy_row<D>(pr, base + (size_t)j * p.ld + p.v_off, acc);
  }
  const float M = warp_max(m);
  const float corr = (m == -INFINITY) ? 0.f : __expf(m - M);
  l = warp_sum(l * corr);
  const float inv_l = 1.0f / l;
  T* orow = reinterpret_cast<T*>(p.out) + ((size_t)b * p.T + i) * p.ld_out + h * D;
  if constexpr (D <= 32) {
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = c < D ? acc[c] * corr : 0.f;
    warp_reduce_scatter32(v, lane);
    if (lane < D) store_elem(orow + lane, v[0] * inv_l);
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float v = warp_sum(acc[c] * corr) * inv_l;
      if (lane == (c & 31)) store_elem(orow + c, v);
    }
  }
  if (lane == 0) p.lse[idx] = M + __logf(l);
}

// dQ (+ delta): one warp per query
template <int D, typename T>
__global__ void __launch_bounds__(kAttnWarps * 32) attn_bwd_q_kernel(gvk_attn_bwd_params bp) {
  const uint64_t seed_eff = salted_seed(bp.f.seed, bp.f.seed_salt);
  const gvk_attn_fwd_params& p = bp.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * kAttnWarps + warp;
  if (idx >= (long long)p.B * p.H * p.T) return;
  const int i = (int)(idx % p.T);
  const int bh = (int)(idx / p.T);
  const int h = bh % p.H, b = bh / p.H;
  const T* base = reinterpret_cast<const T*>(p.qkv) + (size_t)b * p.T * p.ld + h * D;
  float q[D], dO[D];
  load_vec<D>(base + (size_t)i * p.ld + p.q_off, q);
  load_vec<D>(reinterpret_cast<const T*>(bp.dout) + ((size_t)b * p.T + i) * bp.ld_dout + h * D, dO);
  float delta;
  {
    float o[D];
    load_vec<D>(reinterpret_cast<const T*>(p.out) + ((size_t)b * p.T + i) * p.ld_out + h * D, o);
    delta = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) delta = fmaf(o[c], dO[c], delta);
  }
#pragma unroll
  for (int c = 0; c < D; ++c) q[c] *= p.scale;
  const float lse = p.lse[idx];
  const bool windowed = p.win_d > 0;
  Box box;
  int count = p.T;
  if (windowed) {
    box = make_box(p, i, false);
    count = box.count();
  }
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  float dq[D];
#pragma unroll
  for (int c = 0; c < D; ++c) dq[c] = 0.f;
  for (int t = lane; t < count; t += 32) {
    const int j = windowed ? box.token(t, p.grid_h, p.grid_w) : t;
    const T* kr = base + (size_t)j * p.ld + p.k_off;
    const float pr = __expf(dot_row<D>(kr, q) - lse);
    float dp = dot_row<D>(base + (size_t)j * p.ld + p.v_off, dO);
    if (p.drop_p > 0.f) dp *= prob_drop(p, seed_eff, ((uint64_t)bh * p.T + i) * p.T + j, inv_keep);
    const float ds = pr * (dp - delta);
    axpy_row<D>(ds, kr, dq);
  }
  T* drow = reinterpret_cast<T*>(bp.dqkv) + ((size_t)b * p.T + i) * bp.ld_dqkv + h * D + p.q_off;
  if constexpr (D <= 32) {
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = c < D ? dq[c] : 0.f;
    warp_reduce_scatter32(v, lane);
    if (lane < D) store_elem(drow + lane, v[0] * p.scale);
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float v = warp_sum(dq[c]) * p.scale;
      if (lane == (c & 31)) store_elem(drow + c, v);
    }
  }
  if (lane == 0) bp.delta[idx] = delta;
}

// dK, dV: one warp per key
template <int D, typename T>
__global__ void __launch_bounds__(kAttnWarps * 32) attn_bwd_kv_kernel(gvk_attn_bwd_params bp) {
  const uint64_t seed_eff = salted_seed(bp.f.seed, bp.f.seed_salt);
  const gvk_attn_fwd_params& p = bp.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * kAttnWarps + warp;
  if (idx >= (long long)p.B * p.H * p.T) return;
  const int j = (int)(idx % p.T);
  const int bh = (int)(idx / p.T);
  const int h = bh % p.H, b = bh / p.H;
  const T* base = reinterpret_cast<const T*>(p.qkv) + (size_t)b * p.T * p.ld + h * D;
  const T* dbase = reinterpret_cast<const T*>(bp.dout) + (size_t)b * p.T * bp.ld_dout + h * D;
  // this key's (pre-scaled) k and v rows, the same in every lane: registers for the 20/32-wide latents, shared memory for D = 64
  constexpr int DR = D <= 32 ? D : 1;
  __shared__ float skv[D <= 32 ? 1 : kAttnWarps][2][D <= 32 ? 1 : D];
  float ks_r[DR], vs_r[DR];
  const float* ks;
  const float* vs;
  if constexpr (D <= 32) {
    load_vec<D>(base + (size_t)j * p.ld + p.k_off, ks_r);
    load_vec<D>(base + (size_t)j * p.ld + p.v_off, vs_r);
#pragma unroll
    for (int c = 0; c < D; ++c) ks_r[c] *= p.scale;
    ks = ks_r;
    vs = vs_r;
  } else {
    for (int c = lane; c < D; c += 32) {
      skv[warp][0][c] = ld_as_float(base + (size_t)j * p.ld + p.k_off + c) * p.scale;
      skv[warp][1][c] = ld_as_float(base + (size_t)j * p.ld + p.v_off + c);
    }
    __syncwarp();
    ks = skv[warp][0];
    vs = skv[warp][1];
  }
  const bool windowed = p.win_d > 0;
  Box box;
  int count = p.T;
  if (windowed) {
    box = make_box(p, j, true);
    count = box.count();
  }
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  float dk[D], dv[D];
#pragma unroll
  for (int c = 0; c < D; ++c) dk[c] = dv[c] = 0.f;
  for (int t = lane; t < count; t += 32) {
    const int i = windowed ? box.token(t, p.grid_h, p.grid_w) : t;
    const T* qr = base + (size_t)i * p.ld + p.q_off;
    const T* dor = dbase + (size_t)i * bp.ld_dout;
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 4) {
      float qv[4], dv4[4];
      load_vec<4>(qr + c0, qv);
      load_vec<4>(dor + c0, dv4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s = fmaf(qv[u], ks[c0 + u], s);
        dp = fmaf(dv4[u], vs[c0 + u], dp);
      }
    }
    const long long qi = (long long)bh * p.T + i;
    const float pr = __expf(s - p.lse[qi]);
    float mult = 1.f;
    if (p.drop_p > 0.f) mult = prob_drop(p, seed_eff, (uint64_t)qi * p.T + j, inv_keep);
    const float ds = pr * (dp * mult - bp.delta[qi]);
    axpy_row<D>(pr * mult, dor, dv);
    axpy_row<D>(ds, qr, dk);
  }
  T* drow = reinterpret_cast<T*>(bp.dqkv) + ((size_t)b * p.T + j) * bp.ld_dqkv + h * D;
  if constexpr (D <= 32) {
    float v[64];           // interleaved so that lane c ends with dk[c], dv[c]
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      v[2 * c] = c < D ? dk[c] : 0.f;
      v[2 * c + 1] = c < D ? dv[c] : 0.f;
    }
    warp_reduce_scatter64(v, lane);
    if (lane < D) {
      store_elem(drow + p.k_off + lane, v[0] * p.scale);
      store_elem(drow + p.v_off + lane, v[1]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float a = warp_sum(dk[c]) * p.scale;
      const float v = warp_sum(dv[c]);
      if (lane == (c & 31)) {
        store_elem(drow + p.k_off + c, a);
        store_elem(drow + p.v_off + c, v);
      }
    }
  }
}

static int check_attn(const gvk_attn_fwd_params& p, const char* who) {
  GVK_CHECK_ARG(p.qkv && p.out && p.lse, "%s: null pointer", who);
  GVK_CHECK_ARG(p.B > 0 && p.T > 0 && p.H > 0, "%s: bad shape B=%d T=%d H=%d", who, p.B, p.T, p.H);
  GVK_CHECK_ARG(p.D == 20 || p.D == 32 || p.D == 64, "%s: D=%d unsupported (20, 32, 64)", who, p.D);
  GVK_CHECK_ARG(p.dtype == GVK_F32 || p.dtype == GVK_BF16, "%s: bad dtype", who);
  GVK_CHECK_ARG(p.ld % 4 == 0 && p.ld_out % 4 == 0 && p.q_off % 4 == 0 && p.k_off % 4 == 0 && p.v_off % 4 == 0, "%s: ld / offsets must be multiples of 4", who);
  GVK_CHECK_ARG(p.win_d == 0 || (p.win_d > 0 && p.win_h > 0 && p.win_w > 0 && p.grid_d * p.grid_h * p.grid_w == p.T),
                "%s: windowed attention needs T == grid_d*grid_h*grid_w", who);
  GVK_CHECK_ARG(p.drop_p >= 0.f && p.drop_p < 1.f, "%s: drop_p out of range", who);
  return GVK_OK;
}

#define GVK_ATTN_DISPATCH(p, KERNEL, ARG)                                                   \
  do {                                                                                      \
    const long long total = (long long)(p).B * (p).H * (p).T;                               \
    const int grid = (int)((total + kAttnWarps - 1) / kAttnWarps);                          \
    const bool f32 = (p).dtype == GVK_F32;                                                  \
    if ((p).D == 20) {                                                                      \
      if (f32) KERNEL<20, float><<<grid, kAttnWarps * 32, 0, stream>>>(ARG);                \
      else KERNEL<20, __nv_bfloat16><<<grid, kAttnWarps * 32, 0, stream>>>(ARG);            \
    } else if ((p).D == 32) {                                                               \
      if (f32) KERNEL<32, float><<<grid, kAttnWarps * 32, 0, stream>>>(ARG);                \
      else KERNEL<32, __nv_bfloat16><<<grid, kAttnWarps * 32, 0, stream>>>(ARG);            \
    } else {                                                                                \
      if (f32) KERNEL<64, float><<<grid, kAttnWarps * 32, 0, stream>>>(ARG);                \
      else KERNEL<64, __nv_bfloat16><<<grid, kAttnWarps * 32, 0, stream>>>(ARG);            \
    }                                                                                       \
  } while (0)

int attn_simt_fwd(const gvk_attn_fwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p, "gvk_attn_simt_fwd: null params");
  int st = check_attn(*p, "gvk_attn_simt_fwd");
  if (st != GVK_OK) return st;
  if (p->precision == GVK_PREC_TF32 && attn_win_tc_supported(p)) return attn_win_tc_fwd(p, stream);
  GVK_ATTN_DISPATCH(*p, attn_fwd_kernel, *p);
  GVK_CHECK_LAUNCH("attn_simt_fwd");
  return GVK_OK;
}

int attn_simt_bwd(const gvk_attn_bwd_params* bp, cudaStream_t stream) {
  GVK_CHECK_ARG(bp && bp->dout && bp->delta && bp->dqkv, "gvk_attn_simt_bwd: null pointer");
  int st = check_attn(bp->f, "gvk_attn_simt_bwd");
  if (st != GVK_OK) return st;
  GVK_CHECK_ARG(bp->ld_dout % 4 == 0 && bp->ld_dqkv % 4 == 0, "gvk_attn_simt_bwd: ld must be multiples of 4");
  if (bp->f.precision == GVK_PREC_TF32 && attn_win_tc_supported(&bp->f)) return attn_win_tc_bwd(bp, stream);
  GVK_ATTN_DISPATCH(bp->f, attn_bwd_q_kernel, *bp);
  GVK_CHECK_LAUNCH("attn_simt_bwd_q");
  GVK_ATTN_DISPATCH(bp->f, attn_bwd_kv_kernel, *bp);
  GVK_CHECK_LAUNCH("attn_simt_bwd_kv");
  return GVK_OK;
}

}  // namespace gvk
