// gvk_attn_win_tc.cu — GAViKO's window-sparse LocalSelfAttention core (model/gaviko.py:229-244 and its autograd) on the tensor cores:
// mma.sync m16n8k8, tf32 operands (rounded to nearest), fp32 accumulate and softmax.  Selected by precision = GVK_PREC_TF32 in
// gvk_attn_fwd_params (the bf16 compute mode); the exact-fp32 form is gvk_attn_simt.cu.
//
// Why: the one-warp-per-query SIMT kernel re-reads every key / value row of a query's <= 216-key window through L1 (35 KB per query) and
// spends ~70 instructions per probability on Philox; it is bound by L1 wavefronts and issue slots at ~4 % of the fp32 FMA rate.  Here a CTA
// stages the contiguous token range its query chunk can see (K, V rows rounded to tf32, plus one word of one-hot grid coordinates per
// token) in shared memory once, each warp owns 16 queries, and every 8-key tile is three MMAs for S = Q K^T, an online softmax on the
// C fragment, and three MMAs for O += P V.  The window is applied per element with one AND + compare: a query row carries the OR of the
// one-hot ranges it may see, a key its own one-hot coordinates, allowed <=> (row & key) == key.  Key tiles with no allowed element are
// skipped by a warp vote.  The C fragment of S feeds the A operand of the second MMA without a shuffle by permuting the key index inside
// the tile (k slot t <-> key 2t, slot t+4 <-> key 2t+1) on both operands.
//
// MMA rows g / g+8 hold queries 2g / 2g+1 of the warp's 16, so a lane owns aligned 2x2 blocks of the (query, key) matrix in all three
// kernels; one Philox call per block gives the four dropout decisions of the block in the forward, dQ and dK/dV kernels alike.
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {
namespace wtc {

constexpr int kMaxWarps = 8;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr uint32_t kNoToken = 0xFFFFFFFFu;   // coordinate word of a padding row: never a subset of a row mask (bit 31 is unused)

struct Geo {
  int chunk, nchunks;            // tokens per CTA, CTAs per (batch, head)
  int plane;                     // grid_h * grid_w
  int max_stage;                 // shared-memory rows
  int lo_d, hi_d, lo_h, hi_h, lo_w, hi_w;   // key j is visible from query i  <=>  i_ax - lo <= j_ax <= i_ax + hi  on every axis
};

template <int D>
struct Cfg {
  static constexpr int KS = (D + 7) / 8;                 // 8-wide feature steps (k steps of S / dP, n tiles of O / dQ / dK / dV)
  static constexpr int S = (D % 8 == 4) ? D : D + 4;     // row stride = 4 mod 8: both fragment read patterns are bank-conflict free
};

__device__ __forceinline__ uint32_t range_bits(int a, int b, int shift) { return ((2u << b) - (1u << a)) << shift; }   // bits a..b

struct Coord { int d, h, w; };
__device__ __forceinline__ Coord coord_of(const gvk_attn_fwd_params& p, int tok) {
  return {tok / (p.grid_h * p.grid_w), (tok / p.grid_w) % p.grid_h, tok % p.grid_w};
}
__device__ __forceinline__ uint32_t token_bits(const gvk_attn_fwd_params& p, int tok) {
  const Coord c = coord_of(p, tok);
  return (1u << c.d) | (1u << (p.grid_d + c.h)) | (1u << (p.grid_d + p.grid_h + c.w));
}
// Box of the tokens related to `tok`: its keys (transposed = false) or the queries that see it (true).  Returns the row mask and the
// first / last token of the box.
__device__ __forceinline__ uint32_t box_of(const gvk_attn_fwd_params& p, const Geo& g, int tok, bool transposed, int& first, int& last) {
  const Coord c = coord_of(p, tok);
  const int ld = transposed ? g.hi_d : g.lo_d, hd = transposed ? g.lo_d : g.hi_d;
  const int lh = transposed ? g.hi_h : g.lo_h, hh = transposed ? g.lo_h : g.hi_h;
  const int lw = transposed ? g.hi_w : g.lo_w, hw = transposed ? g.lo_w : g.hi_w;
  const int d0 = max(0, c.d - ld), d1 = min(p.grid_d - 1, c.d + hd);
  const int h0 = max(0, c.h - lh), h1 = min(p.grid_h - 1, c.h + hh);
  const int w0 = max(0, c.w - lw), w1 = min(p.grid_w - 1, c.w + hw);
  first = (d0 * p.grid_h + h0) * p.grid_w + w0;
  last = (d1 * p.grid_h + h1) * p.grid_w + w1;
  return range_bits(d0, d1, 0) | range_bits(h0, h1, p.grid_d) | range_bits(w0, w1, p.grid_d + p.grid_h);
}
// Token range [lo, hi) a CTA stages for the chunk [a, b]: whole planes, rounded to the 16-token grid of the tile pairs.
__device__ __forceinline__ void stage_range(const gvk_attn_fwd_params& p, const Geo& g, int a, int b, bool transposed, int& lo, int& hi) {
  const int da = a / g.plane, db = b / g.plane;
  const int back = transposed ? g.hi_d : g.lo_d, fwd = transposed ? g.lo_d : g.hi_d;
  lo = (max(0, da - back) * g.plane) & ~15;
  hi = (min(p.T, (min(p.grid_d - 1, db + fwd) + 1) * g.plane) + 15) & ~15;
}

// dropout multipliers of the aligned 2x2 block (queries 2*ib, 2*ib+1) x (keys 2*jb, 2*jb+1): x (even q, even k) y (even q, odd k)
// z (odd q, even k) w (odd q, odd k)
__device__ __forceinline__ float4 block_drop(const gvk_attn_fwd_params& p, uint64_t seed, int bh, int ib, int jb, float inv_keep) {
  const uint64_t t2 = (uint64_t)(p.T + 1) / 2;
  const uint64_t ctr = p.offset + ((uint64_t)bh * t2 + ib) * t2 + jb;
  const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x77696eu, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return make_float4(u32_to_unit(r.x) >= p.drop_p ? inv_keep : 0.f, u32_to_unit(r.y) >= p.drop_p ? inv_keep : 0.f,
                     u32_to_unit(r.z) >= p.drop_p ? inv_keep : 0.f, u32_to_unit(r.w) >= p.drop_p ? inv_keep : 0.f);
}

// Stage rows [lo, lo + n) of two [*, ld] fp32 matrices (D columns from column `col`) as tf32 into dst[row * S + c]; rows >= T are zero.
// Four 16-byte loads per matrix are in flight per thread.
template <int D>
__device__ __forceinline__ void stage_rows2(float* dst_a, const float* __restrict__ src_a, size_t ld_a, int col_a, float* dst_b,
                                            const float* __restrict__ src_b, size_t ld_b, int col_b, int lo, int n, int T) {
  constexpr int S = Cfg<D>::S, V = D / 4, U = 4;
  const int total = n * V;
  for (int idx0 = threadIdx.x; idx0 < total; idx0 += blockDim.x * U) {
    float4 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = idx0 + u * blockDim.x;
      const int row = idx / V, c4 = idx - row * V;
      va[u] = vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total && lo + row < T) {
        va[u] = *reinterpret_cast<const float4*>(src_a + (size_t)(lo + row) * ld_a + col_a + c4 * 4);
        vb[u] = *reinterpret_cast<const float4*>(src_b + (size_t)(lo + row) * ld_b + col_b + c4 * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = idx0 + u * blockDim.x;
      const int row = idx / V, c4 = idx - row * V;
      if (idx < total) {
        *reinterpret_cast<float4*>(dst_a + row * S + c4 * 4) = make_float4(tf32_round(va[u].x), tf32_round(va[u].y), tf32_round(va[u].z), tf32_round(va[u].w));
        *reinterpret_cast<float4*>(dst_b + row * S + c4 * 4) = make_float4(tf32_round(vb[u].x), tf32_round(vb[u].y), tf32_round(vb[u].z), tf32_round(vb[u].w));
      }
    }
  }
  if (threadIdx.x < 8) dst_a[n * S + threadIdx.x] = dst_b[n * S + threadIdx.x] = 0.f;   // the n tile that overhangs D reads past the last row
}

// A fragments of one 16-row tile (rows r0 + 2g, r0 + 2g + 1) of a [*, ld] matrix, D columns from `col`, times `mul`; rows > last are zero.
template <int D>
__device__ __forceinline__ void load_a(uint32_t (&a)[Cfg<D>::KS][4], const float* __restrict__ src, size_t ld, int col, int r0, int last, float mul,
                                       int g, int t) {
#pragma unroll
  for (int ks = 0; ks < Cfg<D>::KS; ++ks) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = r0 + 2 * g + (e & 1), c = ks * 8 + t + (e >> 1) * 4;
      float v = 0.f;
      if (row <= last && c < D) v = src[(size_t)row * ld + col + c] * mul;
      a[ks][e] = f2tf32(v);
    }
  }
}
// acc(16 x 8) = A(16 x D) * M^T for the 8 staged rows at `rows` (B operand: k = feature, n = staged row)
template <int D>
__device__ __forceinline__ void mma_rows_t(float (&acc)[4], const uint32_t (&a)[Cfg<D>::KS][4], const float* rows, int g, int t) {
  constexpr int S = Cfg<D>::S;
  acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
  const float* r = rows + g * S + t;
#pragma unroll
  for (int ks = 0; ks < Cfg<D>::KS; ++ks) {
    const uint32_t b0 = __float_as_uint(r[ks * 8]);
    const uint32_t b1 = (ks * 8 + 4 < D) ? __float_as_uint(r[ks * 8 + 4]) : 0u;   // D % 8 == 4: the last half step is padding
    mma_tf32(acc, a[ks][0], a[ks][1], a[ks][2], a[ks][3], b0, b1);
  }
}
// acc[nt](16 x 8) += P(16 x 8 staged rows, C-fragment order) * M[rows, nt*8 .. nt*8+7]  (B operand: k slot t <-> row 2t, t+4 <-> row 2t+1)
template <int D>
__device__ __forceinline__ void mma_cfrag(float (&acc)[Cfg<D>::KS][4], const float (&pc)[4], const float* rows, int g, int t) {
  constexpr int S = Cfg<D>::S;
  const uint32_t a0 = f2tf32(pc[0]), a1 = f2tf32(pc[2]), a2 = f2tf32(pc[1]), a3 = f2tf32(pc[3]);
  const float* r = rows + 2 * t * S + g;
#pragma unroll
  for (int nt = 0; nt < Cfg<D>::KS; ++nt) mma_tf32(acc[nt], a0, a1, a2, a3, __float_as_uint(r[nt * 8]), __float_as_uint(r[S + nt * 8]));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
// rows r0 + 2g (c[0], c[1]) and r0 + 2g + 1 (c[2], c[3]) of the n tiles, columns nt*8 + 2t, +1, times mul -> dst[row * ld + col ...]
template <int D>
__device__ __forceinline__ void store_tile(float* dst, size_t ld, int col, const float (&acc)[Cfg<D>::KS][4], int r0, int last, float mul_a, float mul_b,
                                           int g, int t) {
#pragma unroll
  for (int nt = 0; nt < Cfg<D>::KS; ++nt) {
    const int c = nt * 8 + 2 * t;
    if (c < D) {
      if (r0 + 2 * g <= last) *reinterpret_cast<float2*>(dst + (size_t)(r0 + 2 * g) * ld + col + c) = make_float2(acc[nt][0] * mul_a, acc[nt][1] * mul_a);
      if (r0 + 2 * g + 1 <= last) *reinterpret_cast<float2*>(dst + (size_t)(r0 + 2 * g + 1) * ld + col + c) = make_float2(acc[nt][2] * mul_b, acc[nt][3] * mul_b);
    }
  }
}

// Per-warp tile bookkeeping shared by the three kernels: the 16 tokens r0..r0+15 of this warp, their row masks and the token range they touch.
struct WarpTile {
  int r0, last;          // first token of the tile, last valid token of the chunk
  uint32_t mask_a, mask_b;
  int t_first, t_last;   // first / last related token over the tile
};
__device__ __forceinline__ WarpTile warp_tile(const gvk_attn_fwd_params& p, const Geo& g, int r0, int last, bool transposed, int lane) {
  WarpTile w;
  w.r0 = r0;
  w.last = last;
  int first = 0x7fffffff, lastk = -1;
  uint32_t m = 0;
  const int tok = r0 + (lane & 15);
  if (tok <= last) m = box_of(p, g, tok, transposed, first, lastk);
  w.t_first = __reduce_min_sync(0xffffffffu, first);
  w.t_last = __reduce_max_sync(0xffffffffu, lastk);
  const int gq = lane >> 2;
  w.mask_a = __shfl_sync(0xffffffffu, m, 2 * gq);
  w.mask_b = __shfl_sync(0xffffffffu, m, 2 * gq + 1);
  return w;
}

// window test of this lane's 2x2 elements of the 8-token tile at staged row r: bit e set <=> element e of the C fragment is allowed
__device__ __forceinline__ uint32_t tile_allow(const WarpTile& w, const uint32_t* cb, int r, int t) {
  const uint2 c = *reinterpret_cast<const uint2*>(cb + r + 2 * t);
  return ((w.mask_a & c.x) == c.x ? 1u : 0u) | ((w.mask_a & c.y) == c.y ? 2u : 0u) | ((w.mask_b & c.x) == c.x ? 4u : 0u) | ((w.mask_b & c.y) == c.y ? 8u : 0u);
}

// ------------------------------------------------------------------------------------------------------------------------------------
// All three kernels walk the related tokens 16 at a time (two independent 8-token MMA tiles per iteration, for instruction-level parallelism
// and half the shuffles per key).
template <int D>
__global__ void __launch_bounds__(kMaxWarps * 32, 2) win_fwd_kernel(gvk_attn_fwd_params p, Geo geo) {
  const uint64_t seed_eff = salted_seed(p.seed, p.seed_salt);
  constexpr int S = Cfg<D>::S, KS = Cfg<D>::KS;
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = Ks + geo.max_stage * S + 8;
  uint32_t* cb = reinterpret_cast<uint32_t*>(Vs + geo.max_stage * S + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int chunk_id = blockIdx.x % geo.nchunks, bh = blockIdx.x / geo.nchunks;
  const int h = bh % p.H, b = bh / p.H;
  const int a = chunk_id * geo.chunk, last = min(p.T, a + geo.chunk) - 1;
  int slo, shi;
  stage_range(p, geo, a, last, false, slo, shi);
  const float* base = reinterpret_cast<const float*>(p.qkv) + (size_t)b * p.T * p.ld + h * D;
  stage_rows2<D>(Ks, base, p.ld, p.k_off, Vs, base, p.ld, p.v_off, slo, shi - slo, p.T);
  for (int r = threadIdx.x; r < shi - slo; r += blockDim.x) cb[r] = slo + r < p.T ? token_bits(p, slo + r) : kNoToken;
  __syncthreads();
  const int r0 = a + warp * 16;
  if (r0 > last) return;
  const WarpTile w = warp_tile(p, geo, r0, last, false, lane);
  uint32_t qa[KS][4];
  load_a<D>(qa, base, p.ld, p.q_off, r0, last, p.scale * kLog2e, g, t);
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;
  float o[KS][4];
#pragma unroll
  for (int nt = 0; nt < KS; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  for (int j0 = w.t_first & ~15; j0 <= w.t_last; j0 += 16) {
    const int r = j0 - slo;
    const uint32_t al[2] = {tile_allow(w, cb, r, t), tile_allow(w, cb, r + 8, t)};
    if (!__any_sync(0xffffffffu, (al[0] | al[1]) != 0u)) continue;
    float s[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) mma_rows_t<D>(s[u], qa, Ks + (r + 8 * u) * S, g, t);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[u][e] = (al[u] >> e) & 1u ? s[u][e] : -INFINITY;
    }
    const float n_a = fmaxf(m_a, quad_max(fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]))));
    const float n_b = fmaxf(m_b, quad_max(fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]))));
    const float u_a = n_a == -INFINITY ? 0.f : n_a, u_b = n_b == -INFINITY ? 0.f : n_b;
    const float corr_a = fast_ex2(m_a - u_a), corr_b = fast_ex2(m_b - u_b);
    m_a = n_a;
    m_b = n_b;
    float pr[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      pr[u][0] = fast_ex2(s[u][0] - u_a); pr[u][1] = fast_ex2(s[u][1] - u_a); pr[u][2] = fast_ex2(s[u][2] - u_b); pr[u][3] = fast_ex2(s[u][3] - u_b);
    }
    l_a = l_a * corr_a + (pr[0][0] + pr[0][1]) + (pr[1][0] + pr[1][1]);
    l_b = l_b * corr_b + (pr[0][2] + pr[0][3]) + (pr[1][2] + pr[1][3]);
#pragma unroll
    for (int nt = 0; nt < KS; ++nt) {
      o[nt][0] *= corr_a; o[nt][1] *= corr_a; o[nt][2] *= corr_b; o[nt][3] *= corr_b;
    }
    if (p.drop_p > 0.f) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 d = block_drop(p, seed_eff, bh, (r0 >> 1) + g, ((j0 + 8 * u) >> 1) + t, inv_keep);
        pr[u][0] *= d.x; pr[u][1] *= d.y; pr[u][2] *= d.z; pr[u][3] *= d.w;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) mma_cfrag<D>(o, pr[u], Vs + (r + 8 * u) * S, g, t);
  }
  l_a = quad_sum(l_a);
  l_b = quad_sum(l_b);
  float* out = reinterpret_cast<float*>(p.out) + (size_t)b * p.T * p.ld_out + h * D;
  store_tile<D>(out, p.ld_out, 0, o, r0, last, 1.0f / l_a, 1.0f / l_b, g, t);
  if (t == 0) {
    if (r0 + 2 * g <= last) p.lse[(size_t)bh * p.T + r0 + 2 * g] = (m_a + __log2f(l_a)) * kLn2;
    if (r0 + 2 * g + 1 <= last) p.lse[(size_t)bh * p.T + r0 + 2 * g + 1] = (m_b + __log2f(l_b)) * kLn2;
  }
}

// dQ and delta = rowsum(O * dO): same tiling as the forward
template <int D>
__global__ void __launch_bounds__(kMaxWarps * 32, 2) win_dq_kernel(gvk_attn_bwd_params bp, Geo geo) {
  const uint64_t seed_eff = salted_seed(bp.f.seed, bp.f.seed_salt);
  constexpr int S = Cfg<D>::S, KS = Cfg<D>::KS;
  const gvk_attn_fwd_params& p = bp.f;
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = Ks + geo.max_stage * S + 8;
  uint32_t* cb = reinterpret_cast<uint32_t*>(Vs + geo.max_stage * S + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int chunk_id = blockIdx.x % geo.nchunks, bh = blockIdx.x / geo.nchunks;
  const int h = bh % p.H, b = bh / p.H;
  const int a = chunk_id * geo.chunk, last = min(p.T, a + geo.chunk) - 1;
  int slo, shi;
  stage_range(p, geo, a, last, false, slo, shi);
  const float* base = reinterpret_cast<const float*>(p.qkv) + (size_t)b * p.T * p.ld + h * D;
  stage_rows2<D>(Ks, base, p.ld, p.k_off, Vs, base, p.ld, p.v_off, slo, shi - slo, p.T);
  for (int r = threadIdx.x; r < shi - slo; r += blockDim.x) cb[r] = slo + r < p.T ? token_bits(p, slo + r) : kNoToken;
  __syncthreads();
  const int r0 = a + warp * 16;
  if (r0 > last) return;
  const WarpTile w = warp_tile(p, geo, r0, last, false, lane);
  const float* dob = reinterpret_cast<const float*>(bp.dout) + (size_t)b * p.T * bp.ld_dout + h * D;
  const float* ob = reinterpret_cast<const float*>(p.out) + (size_t)b * p.T * p.ld_out + h * D;
  uint32_t qa[KS][4], da[KS][4];
  load_a<D>(qa, base, p.ld, p.q_off, r0, last, p.scale * kLog2e, g, t);
  load_a<D>(da, dob, bp.ld_dout, 0, r0, last, 1.f, g, t);
  const int row_a = r0 + 2 * g, row_b = row_a + 1;
  float delta_a = 0.f, delta_b = 0.f, lse_a = 0.f, lse_b = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = ks * 8 + t + e * 4;
      if (c < D) {
        if (row_a <= last) delta_a = fmaf(ob[(size_t)row_a * p.ld_out + c], dob[(size_t)row_a * bp.ld_dout + c], delta_a);
        if (row_b <= last) delta_b = fmaf(ob[(size_t)row_b * p.ld_out + c], dob[(size_t)row_b * bp.ld_dout + c], delta_b);
      }
    }
  }
  delta_a = quad_sum(delta_a);
  delta_b = quad_sum(delta_b);
  if (row_a <= last) lse_a = p.lse[(size_t)bh * p.T + row_a] * kLog2e;
  if (row_b <= last) lse_b = p.lse[(size_t)bh * p.T + row_b] * kLog2e;
  if (t == 0) {
    if (row_a <= last) bp.delta[(size_t)bh * p.T + row_a] = delta_a;
    if (row_b <= last) bp.delta[(size_t)bh * p.T + row_b] = delta_b;
  }
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  float dq[KS][4];
#pragma unroll
  for (int nt = 0; nt < KS; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
  for (int j0 = w.t_first & ~15; j0 <= w.t_last; j0 += 16) {
    const int r = j0 - slo;
    const uint32_t al[2] = {tile_allow(w, cb, r, t), tile_allow(w, cb, r + 8, t)};
    if (!__any_sync(0xffffffffu, (al[0] | al[1]) != 0u)) continue;
    float s[2][4], dp[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      mma_rows_t<D>(s[u], qa, Ks + (r + 8 * u) * S, g, t);
      mma_rows_t<D>(dp[u], da, Vs + (r + 8 * u) * S, g, t);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (p.drop_p > 0.f) {
        const float4 d = block_drop(p, seed_eff, bh, (r0 >> 1) + g, ((j0 + 8 * u) >> 1) + t, inv_keep);
        dp[u][0] *= d.x; dp[u][1] *= d.y; dp[u][2] *= d.z; dp[u][3] *= d.w;
      }
      float ds[4];
      ds[0] = al[u] & 1u ? fast_ex2(s[u][0] - lse_a) * (dp[u][0] - delta_a) : 0.f;
      ds[1] = al[u] & 2u ? fast_ex2(s[u][1] - lse_a) * (dp[u][1] - delta_a) : 0.f;
      ds[2] = al[u] & 4u ? fast_ex2(s[u][2] - lse_b) * (dp[u][2] - delta_b) : 0.f;
      ds[3] = al[u] & 8u ? fast_ex2(s[u][3] - lse_b) * (dp[u][3] - delta_b) : 0.f;
      mma_cfrag<D>(dq, ds, Ks + (r + 8 * u) * S, g, t);
    }
  }
  float* dst = reinterpret_cast<float*>(bp.dqkv) + (size_t)b * p.T * bp.ld_dqkv + h * D;
  store_tile<D>(dst, bp.ld_dqkv, p.q_off, dq, r0, last, p.scale, p.scale, g, t);
}

// dK, dV: a warp owns 16 keys, the tiles run over the queries that see them (the transposed window is again a box); no atomics
template <int D>
__global__ void __launch_bounds__(kMaxWarps * 32, 2) win_dkv_kernel(gvk_attn_bwd_params bp, Geo geo) {
  const uint64_t seed_eff = salted_seed(bp.f.seed, bp.f.seed_salt);
  constexpr int S = Cfg<D>::S, KS = Cfg<D>::KS;
  const gvk_attn_fwd_params& p = bp.f;
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;
  float* Os = Qs + geo.max_stage * S + 8;                  // dO rows
  float* lse2 = Os + geo.max_stage * S + 8;
  float* dlt = lse2 + geo.max_stage;
  uint32_t* cb = reinterpret_cast<uint32_t*>(dlt + geo.max_stage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int chunk_id = blockIdx.x % geo.nchunks, bh = blockIdx.x / geo.nchunks;
  const int h = bh % p.H, b = bh / p.H;
  const int a = chunk_id * geo.chunk, last = min(p.T, a + geo.chunk) - 1;
  int slo, shi;
  stage_range(p, geo, a, last, true, slo, shi);
  const float* base = reinterpret_cast<const float*>(p.qkv) + (size_t)b * p.T * p.ld + h * D;
  const float* dob = reinterpret_cast<const float*>(bp.dout) + (size_t)b * p.T * bp.ld_dout + h * D;
  stage_rows2<D>(Qs, base, p.ld, p.q_off, Os, dob, bp.ld_dout, 0, slo, shi - slo, p.T);
  for (int r = threadIdx.x; r < shi - slo; r += blockDim.x) {
    const bool ok = slo + r < p.T;
    cb[r] = ok ? token_bits(p, slo + r) : kNoToken;
    lse2[r] = ok ? p.lse[(size_t)bh * p.T + slo + r] * kLog2e : 0.f;
    dlt[r] = ok ? bp.delta[(size_t)bh * p.T + slo + r] : 0.f;
  }
  __syncthreads();
  const int r0 = a + warp * 16;
  if (r0 > last) return;
  const WarpTile w = warp_tile(p, geo, r0, last, true, lane);
  uint32_t ka[KS][4], va[KS][4];
  load_a<D>(ka, base, p.ld, p.k_off, r0, last, p.scale * kLog2e, g, t);
  load_a<D>(va, base, p.ld, p.v_off, r0, last, 1.f, g, t);
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.f;
  float dk[KS][4], dv[KS][4];
#pragma unroll
  for (int nt = 0; nt < KS; ++nt) dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
  for (int i0 = w.t_first & ~15; i0 <= w.t_last; i0 += 16) {
    const int r = i0 - slo;
    // rows are keys (mask_a: key 2g, mask_b: key 2g+1), columns are queries 2t, 2t+1
    const uint32_t al[2] = {tile_allow(w, cb, r, t), tile_allow(w, cb, r + 8, t)};
    if (!__any_sync(0xffffffffu, (al[0] | al[1]) != 0u)) continue;
    float s[2][4], dp[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      mma_rows_t<D>(s[u], ka, Qs + (r + 8 * u) * S, g, t);
      mma_rows_t<D>(dp[u], va, Os + (r + 8 * u) * S, g, t);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ru = r + 8 * u;
      const float2 ls = *reinterpret_cast<const float2*>(lse2 + ru + 2 * t);
      const float2 dl = *reinterpret_cast<const float2*>(dlt + ru + 2 * t);
      float pd[4] = {al[u] & 1u ? fast_ex2(s[u][0] - ls.x) : 0.f, al[u] & 2u ? fast_ex2(s[u][1] - ls.y) : 0.f,
                     al[u] & 4u ? fast_ex2(s[u][2] - ls.x) : 0.f, al[u] & 8u ? fast_ex2(s[u][3] - ls.y) : 0.f};
      float4 d = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.drop_p > 0.f) d = block_drop(p, seed_eff, bh, ((i0 + 8 * u) >> 1) + t, (r0 >> 1) + g, inv_keep);
      // block_drop: x (even q, even k) y (even q, odd k) z (odd q, even k) w (odd q, odd k);  here element 0 = (key 2g, query 2t),
      // 1 = (key 2g, query 2t+1), 2 = (key 2g+1, query 2t), 3 = (key 2g+1, query 2t+1)
      const float mult[4] = {d.x, d.z, d.y, d.w};
      float ds[4];
      ds[0] = pd[0] * (dp[u][0] * mult[0] - dl.x);
      ds[1] = pd[1] * (dp[u][1] * mult[1] - dl.y);
      ds[2] = pd[2] * (dp[u][2] * mult[2] - dl.x);
      ds[3] = pd[3] * (dp[u][3] * mult[3] - dl.y);
#pragma unroll
      for (int e = 0; e < 4; ++e) pd[e] *= mult[e];
      mma_cfrag<D>(dv, pd, Os + ru * S, g, t);
      mma_cfrag<D>(dk, ds, Qs + ru * S, g, t);
    }
  }
  float* dst = reinterpret_cast<float*>(bp.dqkv) + (size_t)b * p.T * bp.ld_dqkv + h * D;
  store_tile<D>(dst, bp.ld_dqkv, p.k_off, dk, r0, last, p.scale, p.scale, g, t);
  store_tile<D>(dst, bp.ld_dqkv, p.v_off, dv, r0, last, 1.f, 1.f, g, t);
}

// ------------------------------------------------------------------------------------------------------------------------------------
constexpr size_t kSmemLimit = 227 * 1024;

template <int D>
static size_t smem_fwd(int rows) { return ((size_t)2 * (rows * Cfg<D>::S + 8) + rows) * 4; }
template <int D>
static size_t smem_dkv(int rows) { return ((size_t)2 * (rows * Cfg<D>::S + 8) + 3 * rows) * 4; }

static bool make_geo(const gvk_attn_fwd_params& p, Geo& g) {
  if (p.win_d <= 0 || p.dtype != GVK_F32 || (p.D != 20 && p.D != 32)) return false;
  if (p.grid_d + p.grid_h + p.grid_w > 31) return false;
  g.plane = p.grid_h * p.grid_w;
  g.lo_d = p.win_d / 2; g.hi_d = p.win_d - 1 - g.lo_d;
  g.lo_h = p.win_h / 2; g.hi_h = p.win_h - 1 - g.lo_h;
  g.lo_w = p.win_w / 2; g.hi_w = p.win_w - 1 - g.lo_w;
  int chunk = g.plane <= kMaxWarps * 16 ? g.plane * ((kMaxWarps * 16) / g.plane) : 64;
  chunk = std::min(chunk, p.T);
  if (chunk > 1) chunk &= ~1;            // chunks start on even tokens: the 2x2 dropout blocks must line up in all three kernels
  g.chunk = chunk;
  g.nchunks = (p.T + chunk - 1) / chunk;
  // rows a CTA stages: whole planes from (first plane of the chunk - back) to (last plane + forward), both ends rounded to the 16-token
  // grid; back / forward are (lo_d, hi_d) for the forward / dQ kernels and swapped for dK / dV
  int planes = 0;
  for (int c = 0; c < g.nchunks; ++c) {
    const int a = c * chunk, b = std::min(p.T, a + chunk) - 1;
    const int da = a / g.plane, db = b / g.plane;
    planes = std::max(planes, std::min(p.grid_d - 1, db + g.hi_d) - std::max(0, da - g.lo_d) + 1);
    planes = std::max(planes, std::min(p.grid_d - 1, db + g.lo_d) - std::max(0, da - g.hi_d) + 1);
  }
  g.max_stage = std::min((p.T + 15) / 16 * 16, (planes * g.plane + 30) / 16 * 16);
  const size_t need = p.D == 20 ? smem_dkv<20>(g.max_stage) : smem_dkv<32>(g.max_stage);
  return need <= kSmemLimit;
}

template <typename K, typename P>
static int launch(K kernel, const P& params, const Geo& g, const gvk_attn_fwd_params& f, size_t smem, cudaStream_t stream, const char* what) {
  if (smem > 48 * 1024) {
    const int st = cuda_status(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what);
    if (st != GVK_OK) return st;
  }
  const int warps = std::min(kMaxWarps, (g.chunk + 15) / 16);
  kernel<<<g.nchunks * f.B * f.H, warps * 32, smem, stream>>>(params, g);
  GVK_CHECK_LAUNCH(what);
  return GVK_OK;
}

}  // namespace wtc

bool attn_win_tc_supported(const gvk_attn_fwd_params* p) {
  wtc::Geo g;
  return wtc::make_geo(*p, g);
}

int attn_win_tc_fwd(const gvk_attn_fwd_params* p, cudaStream_t stream) {
  wtc::Geo g;
  GVK_CHECK_ARG(wtc::make_geo(*p, g), "gvk_attn_simt_fwd: tensor-core window path does not support this problem");
  if (p->D == 20) return wtc::launch(wtc::win_fwd_kernel<20>, *p, g, *p, wtc::smem_fwd<20>(g.max_stage), stream, "attn_win_tc_fwd");
  return wtc::launch(wtc::win_fwd_kernel<32>, *p, g, *p, wtc::smem_fwd<32>(g.max_stage), stream, "attn_win_tc_fwd");
}

int attn_win_tc_bwd(const gvk_attn_bwd_params* bp, cudaStream_t stream) {
  wtc::Geo g;
  GVK_CHECK_ARG(wtc::make_geo(bp->f, g), "gvk_attn_simt_bwd: tensor-core window path does not support this problem");
  int st;
  if (bp->f.D == 20) {
    st = wtc::launch(wtc::win_dq_kernel<20>, *bp, g, bp->f, wtc::smem_fwd<20>(g.max_stage), stream, "attn_win_tc_dq");
    if (st != GVK_OK) return st;
    return wtc::launch(wtc::win_dkv_kernel<20>, *bp, g, bp->f, wtc::smem_dkv<20>(g.max_stage), stream, "attn_win_tc_dkv");
  }
  st = wtc::launch(wtc::win_dq_kernel<32>, *bp, g, bp->f, wtc::smem_fwd<32>(g.max_stage), stream, "attn_win_tc_dq");
  if (st != GVK_OK) return st;
  return wtc::launch(wtc::win_dkv_kernel<32>, *bp, g, bp->f, wtc::smem_dkv<32>(g.max_stage), stream, "attn_win_tc_dkv");
}

}  // namespace gvk
