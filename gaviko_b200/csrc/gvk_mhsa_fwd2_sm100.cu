// gvk_mhsa_fwd2_sm100.cu — flash-attention FORWARD for the frozen MHSA core, second generation (head dim 64, bf16 operands, fp32 softmax).
//
// Same roles as gvk_mhsa_ws_sm100.cu (one persistent CTA per SM; warp 0 TMA, warps 1 / 2 MMA issuers of tile A / B, warp 3 TMEM, warps 4-7 /
// 8-11 softmax groups of tile A / B, thread = query row), but built on what the backward's timeline traces showed (profiles/mhsa_*_timeline*):
//   * 128-key steps: S = Q K^T is ONE N = 128 MMA group per step (64 clk per instruction for twice the work of an N = 64 one) and every
//     per-step cost of a softmax warp — barrier waits, tcgen05 fences, TMEM load / store round trips — is paid half as often per score;
//   * P has its own TMEM columns, so S(j+1) is issued as soon as the group has pulled S(j) into registers (S_FREE), not after P V(j):
//       TMEM per tile X (256 columns):  S [0,128)   P (bf16 pairs) [128,192)   O [192,256)
//   * barrier waits are polled by one lane per warp, operand descriptors are stepped with one 32-bit add;
//   * the 9-key tail of T = 1033 is an N = 16 / K = 16 step, not a padded 64-key tile.
// The two groups are independent (each owns a tile), so one group's exp2 phase fills the SFU while the other loads / stores / waits.
// Optional Philox dropout on the probabilities (gvk.h, gvk_mhsa_fwd_params).
//
// Measured (B = 64, T = 1033, H = 12): 440 us with the token, 481 us without, against 420-428 us of the 64-key-step kernel, which therefore
// stays the default (GVK_MHSA_IMPL=3 selects this one).  The timeline of CTA 0 (GVK_PIPE_DBG=32, tools/mhsa_trace.py) says why: a softmax
// warp that has the SFU to itself issues one ex2 per ~12 clk (128 exponentials in ~1545 clk; the SFU's rate is 8 clk per warp instruction,
// reached only when two warps of a sub-partition interleave), and there are exactly two softmax warps per sub-partition.  So either the two
// groups run their exp2 phases together (full SFU rate, but the SFU idles during their load / max / store / wait phases: 3430 clk per
// 128-key step of both tiles) or alternately (no idle phase, but 2/3 of the rate: 3016 clk).  Moving a third of the exponentials to an
// FMA-pipe polynomial (Cody-Waite + degree 3, 1e-4 accurate) was also measured, in the backward: slower (1207 vs 1077 us), the extra ~5
// instructions per exponential cost more issue slots than the SFU time they save.  More softmax warps per sub-partition (a 16-warp
// layout at <= 96 registers per thread) is the open direction.
//
// Replaces model/vision_transformer.py:65-71 (softmax(q k^T * scale) v on the prompt-extended sequence); any T (tails masked in-kernel).
#include <algorithm>
#include <cstdlib>

#include "gvk_common.cuh"

namespace gvk {

namespace f2 {
constexpr int kThreads = 384;
constexpr int kD = 64;
constexpr int kTile = 128;                    // query rows per tile, keys per step
constexpr int kTileBytes = kTile * kD * 2;    // 16 KB
constexpr int kStages = 4;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;     // log2 units
constexpr int kColP = 128, kColO = 192, kTileCols = 256;
constexpr uint32_t kDescHi = 0x40004040u;     // SBO 1024 B | version 1 | SWIZZLE_128B

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}
// The SFU (16 ex2 / clk / SM) is the bound of this kernel, and left alone the two softmax groups fall into lockstep (they wait for the same
// K / V stages), so both sit in their exp2 phase together at half rate each and the SFU idles while both load / store / wait.  A token
// (two named barriers, 128 arriving + 128 waiting threads each) makes the exp2 phases mutually exclusive: one group runs its
// exponentials at the full SFU rate while the other does everything else.
__device__ __forceinline__ void token_wait(int id) { asm volatile("barrier.cta.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void token_pass(int id) { asm volatile("barrier.cta.arrive %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (512u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi));
  return d;
}

struct Args {
  int B, T, H, dim;
  int nb, ntail;       // 128-key steps per sequence; valid keys of the last one rounded up to 16
  float scale;
  __nv_bfloat16* out;
  int ld_out;
  float* lse;
  int pairs, num_items;
  MhsaDrop drop;
  uint32_t* trace;
  int dbg;          // GVK_PIPE_DBG & 32: record the timeline of CTA 0 (tools/mhsa_trace.py)
};

enum { BAR_Q_FULL = 0, BAR_Q_EMPTY = 1, BAR_KV_FULL = 2, BAR_KV_EMPTY = BAR_KV_FULL + kStages, BAR_S_FULL = BAR_KV_EMPTY + kStages /*[X]*/, BAR_S_FREE = BAR_S_FULL + 2 /*[X]*/,
       BAR_P_FULL = BAR_S_FREE + 2 /*[X]*/, BAR_PV_DONE = BAR_P_FULL + 2 /*[X]*/, BAR_COUNT = BAR_PV_DONE + 2 };
constexpr int kSmem = 2 * kTileBytes + 2 * kStages * kTileBytes + BAR_COUNT * 8 + 64 + 1024;

template <bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
mhsa_fwd2_kernel(const __grid_constant__ CUtensorMap tma_qkv, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // [2][16 KB]
  uint8_t* sK = sQ + 2 * kTileBytes;               // [kStages][16 KB]
  uint8_t* sV = sK + kStages * kTileBytes;         // [kStages][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStages * kTileBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, dim = a.dim, nb = a.nb;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    for (int i = 0; i < BAR_COUNT; ++i) {
      int count = 1;
      if ((i >= BAR_S_FREE && i < BAR_S_FREE + 2) || (i >= BAR_P_FULL && i < BAR_P_FULL + 2)) count = 4;     // one arrival per softmax warp of the tile
      if (i == BAR_Q_EMPTY || (i >= BAR_KV_EMPTY && i < BAR_KV_EMPTY + kStages)) count = 2;                  // one per MMA issuer
      mbar_init(&bars[i], count);
    }
    fence_barrier_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    reg_dealloc<72>();
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      uint32_t kv_iter = 0, work = 0;
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work) {
        const int bh = item / a.pairs, qp = item - bh * a.pairs;
        const int h = bh % a.H, b = bh / a.H;
        const int q0 = qp * 2 * kTile;
        const bool activeB = q0 + kTile < T;
        mbar_wait(&bars[BAR_Q_EMPTY], (work & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[BAR_Q_FULL], activeB ? 2 * kTileBytes : kTileBytes);
        tma_load_3d(sQ, &tma_qkv, &bars[BAR_Q_FULL], h * kD, q0, b);
        if (activeB) tma_load_3d(sQ + kTileBytes, &tma_qkv, &bars[BAR_Q_FULL], h * kD, q0 + kTile, b);
        for (int j = 0; j < nb; ++j, ++kv_iter) {
          const int st = kv_iter % kStages;
          mbar_wait(&bars[BAR_KV_EMPTY + st], ((kv_iter / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[BAR_KV_FULL + st], 2 * kTileBytes);
          tma_load_3d(sK + st * kTileBytes, &tma_qkv, &bars[BAR_KV_FULL + st], dim + h * kD, j * kTile, b);
          tma_load_3d(sV + st * kTileBytes, &tma_qkv, &bars[BAR_KV_FULL + st], 2 * dim + h * kD, j * kTile, b);
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: warp 1 drives tile A, warp 2 tile B (warp-uniform control flow) =====================
      const int X = warp - 1;
      const uint32_t q_lo = desc_lo_k(smem_u32(sQ + X * kTileBytes));
      const uint32_t tS = tmem + X * kTileCols, tP = tS + kColP, tO = tS + kColO;
      const uint32_t idesc_s_full = make_idesc_bf16(128, kTile, 0, 0), idesc_s_tail = make_idesc_bf16(128, a.ntail, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      uint32_t kv_base = 0, work = 0, g = 0;       // g = score tiles of this issuer so far (phase of the per-tile barriers)
      auto issue_s = [&](int j) {                  // S(j) = Q K_j^T once the group holds S(j-1) in registers
        const uint32_t it = kv_base + j, gg = g + j;
        const int st = it % kStages;
        mbar_wait(&bars[BAR_KV_FULL + st], (it / kStages) & 1);
        if (gg > 0) mbar_wait(&bars[BAR_S_FREE + X], (gg - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t k_lo = desc_lo_k(smem_u32(sK + st * kTileBytes));
          const uint32_t idesc = j == nb - 1 ? idesc_s_tail : idesc_s_full;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS, desc64(q_lo + 2 * k), desc64(k_lo + 2 * k), idesc, k > 0 ? 1u : 0u);
          umma_commit(&bars[BAR_S_FULL + X]);
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, kv_base += nb) {
        const int qp = item % a.pairs;
        const bool activeB = qp * 2 * kTile + kTile < T;
        const int releases = (X == 0 && !activeB) ? 2 : 1;
        // both issuers observe EVERY Q_FULL phase, even for an item tile B sits out (a skipped phase would alias with the one before it)
        mbar_wait(&bars[BAR_Q_FULL], work & 1);
        if (X == 1 && !activeB) continue;          // warp 1 then releases the shared stages for both
        issue_s(0);
        for (int j = 0; j < nb; ++j) {
          const uint32_t it = kv_base + j, gg = g + j;
          const int st = it % kStages;
          if (j + 1 < nb) issue_s(j + 1);
          else if (elect_one())
            for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_Q_EMPTY]);     // the last S MMAs of the item were issued: Q is free when they complete
          __syncwarp();
          mbar_wait(&bars[BAR_P_FULL + X], gg & 1);
          tc_fence_after();
          if (elect_one()) {                       // O (+)= P V_j, P read from TMEM (8 columns per 16 keys), V_j as the MN-major operand
            const uint32_t v_lo = desc_lo_mn(smem_u32(sV + st * kTileBytes));
            const int ksteps = (j == nb - 1 ? a.ntail : kTile) / 16;
            for (int k = 0; k < ksteps; ++k) umma_bf16_ts(tO, tP + 8 * k, desc64(v_lo + 128 * k), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bars[BAR_PV_DONE + X]);
            for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_KV_EMPTY + st]);   // K_j was read by S(j), V_j by P V(j)
          }
          __syncwarp();
        }
        g += nb;
      }
    }
  } else {
    // ===================== softmax groups: thread = query row of its tile =====================
    reg_alloc<208>();
    const int X = (warp - 4) >> 2;
    const int r = (tid - 128) & 127;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + X * kTileCols + lane_off, tP = tS + kColP, tO = tS + kColO;
    const float c2 = a.scale * kLog2e;
    uint32_t g = 0;
    const MhsaDrop drop = kDrop ? mhsa_salted(a.drop) : a.drop;
    Tracer tr; tr.init(a.trace, 1 + X, (a.dbg & 32) && (warp & 3) == 0 && lane == 0);
    const bool use_token = !(a.dbg & 64);
    if (use_token && X == 1) token_pass(1);      // group A holds the token first (barrier 1 + X = "group X may run its exponentials")
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
      const int bh = item / a.pairs, qp = item - bh * a.pairs;
      const int h = bh % a.H, b = bh / a.H;
      const int q0 = qp * 2 * kTile + X * kTile;
      if (q0 >= T) continue;                  // tile B of the last pair may be empty (never for tile A)
      const bool paired = use_token && qp * 2 * kTile + kTile < T;     // a lone tile A has the SFU to itself
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < nb; ++j, ++g) {
        const int ncols = j == nb - 1 ? a.ntail : kTile;      // columns the MMA wrote (multiple of 16)
        const int valid = min(kTile, T - j * kTile);          // columns that are real keys
        tr(0x100 + j);
        mbar_wait_warp(&bars[BAR_S_FULL + X], g & 1, lane);
        tc_fence_after();
        tr(0x200 + j);
        float s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (32 * c < ncols) tmem_ld_32x32(tS + 32 * c, *reinterpret_cast<float(*)[32]>(&s[32 * c]));
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_S_FREE + X]);      // S(j) is in registers: S(j+1) may overwrite it
        tr(0x300 + j);
        if (valid < kTile) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) s[i] = -INFINITY;                   // also covers the columns the tail MMA never wrote
        }
        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]); mx2 = fmaxf(mx2, s[i + 2]); mx3 = fmaxf(mx3, s[i + 3]);
        }
        const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
        const bool resc = (m_new - m_used) * c2 > kRescaleThreshold;   // also true on the first step (m_used = -inf)
        float alpha = 1.f;
        if (resc) {
          alpha = fast_ex2((m_used - m_new) * c2);
          m_used = m_new;
        }
        const float mc = m_used * c2;
        const float2 c2v = make_float2(c2, c2), mcv = make_float2(-mc, -mc);
        // P V(j-1) must be complete before P(j) overwrites its operand and before O is rescaled (it finished long ago: it was issued when the
        // previous step ended)
        if (j > 0) {
          mbar_wait_warp(&bars[BAR_PV_DONE + X], (g - 1) & 1, lane);
          tc_fence_after();
        }
        if (paired) token_wait(1 + X);
        tr(0x400 + j);
        float2 rs01 = make_float2(0.f, 0.f), rs23 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {           // 32 scores -> 16 packed registers -> P columns 16c .. 16c+15
          if (32 * c >= ncols) break;
          uint32_t pk[16];
          uint32_t keep = 0xFFFFFFFFu;          // dropout decisions of these 32 keys (the row sum l stays that of the un-dropped softmax)
          if (kDrop) keep = mhsa_keep16(drop, bh, q0 + r, 8 * j + 2 * c) | (mhsa_keep16(drop, bh, q0 + r, 8 * j + 2 * c + 1) << 16);
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float2 pa = ffma2(make_float2(s[32 * c + 2 * i], s[32 * c + 2 * i + 1]), c2v, mcv);
            float2 pb = ffma2(make_float2(s[32 * c + 2 * i + 2], s[32 * c + 2 * i + 3]), c2v, mcv);
            pa.x = fast_ex2(pa.x); pa.y = fast_ex2(pa.y); pb.x = fast_ex2(pb.x); pb.y = fast_ex2(pb.y);
            rs01 = fadd2(rs01, pa);
            rs23 = fadd2(rs23, pb);
            if (kDrop) {
              if (!((keep >> (2 * i)) & 1u)) pa.x = 0.f;
              if (!((keep >> (2 * i + 1)) & 1u)) pa.y = 0.f;
              if (!((keep >> (2 * i + 2)) & 1u)) pb.x = 0.f;
              if (!((keep >> (2 * i + 3)) & 1u)) pb.y = 0.f;
            }
            pk[i] = pack_bf16x2(pa.x, pa.y);
            pk[i + 1] = pack_bf16x2(pb.x, pb.y);
          }
          tmem_st_32x16(tP + 16 * c, pk);
        }
        l = fmaf(l, alpha, (rs01.x + rs01.y) + (rs23.x + rs23.y));
        if (paired) token_pass(2 - X);
        tr(0x500 + j);
        if (j > 0 && __any_sync(0xffffffffu, resc)) {
          // O (accumulated by the previous steps) must be in the units of the new maximum before P V(j) adds to it
          float o[64];
          tmem_ld_32x32(tO, *reinterpret_cast<float(*)[32]>(&o[0]));
          tmem_ld_32x32(tO + 32, *reinterpret_cast<float(*)[32]>(&o[32]));
          tc_wait_ld();
          const float2 av = make_float2(alpha, alpha);
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            const float2 v = fmul2(make_float2(o[i], o[i + 1]), av);
            o[i] = v.x;
            o[i + 1] = v.y;
          }
          tmem_st_32x32(tO, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
          tmem_st_32x32(tO + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_P_FULL + X]);
        tr(0x600 + j);
      }
      // ---- epilogue: O / l -> bf16, lse
      mbar_wait_warp(&bars[BAR_PV_DONE + X], (g - 1) & 1, lane);
      tc_fence_after();
      float o[64];
      tmem_ld_32x32(tO, *reinterpret_cast<float(*)[32]>(&o[0]));
      tmem_ld_32x32(tO + 32, *reinterpret_cast<float(*)[32]>(&o[32]));
      tc_wait_ld();
      const int row = q0 + r;
      if (row < T) {
        const float inv = (kDrop ? a.drop.inv_keep : 1.0f) / l;
        uint4* dst = reinterpret_cast<uint4*>(a.out + ((size_t)b * T + row) * a.ld_out + h * kD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 pk;
          pk.x = pack_bf16x2(o[8 * c + 0] * inv, o[8 * c + 1] * inv);
          pk.y = pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
          pk.z = pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
          pk.w = pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
          dst[c] = pk;
        }
        a.lse[(size_t)bh * T + row] = m_used * a.scale + __logf(l);
      }
      // the next item's P V(0) (accumulate = 0) is gated by this group's next P_FULL arrival, i.e. after these O reads: no extra barrier
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}

}  // namespace f2

int mhsa_fwd2(const gvk_mhsa_fwd_params* p, cudaStream_t stream) {
  using namespace f2;
  static bool configured = false;
  if (!configured) {
    int st = cuda_status(cudaFuncSetAttribute(mhsa_fwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem), "mhsa_fwd2 smem");
    if (st != GVK_OK) return st;
    st = cuda_status(cudaFuncSetAttribute(mhsa_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem), "mhsa_fwd2 (dropout) smem");
    if (st != GVK_OK) return st;
    configured = true;
  }
  const int dim = p->H * kD;
  CUtensorMap tqkv;
  int st = make_tma_3d_bf16(&tqkv, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, kTile, kD);
  if (st != GVK_OK) return st;
  Args a;
  a.B = p->B; a.T = p->T; a.H = p->H; a.dim = dim; a.scale = p->scale;
  a.nb = (p->T + kTile - 1) / kTile;
  a.ntail = (p->T - (a.nb - 1) * kTile + 15) / 16 * 16;
  a.out = reinterpret_cast<__nv_bfloat16*>(p->out);
  a.ld_out = p->ld_out;
  a.lse = p->lse;
  a.pairs = (a.nb + 1) / 2;
  a.num_items = p->B * p->H * a.pairs;
  a.drop = make_mhsa_drop(p->drop_p, p->seed, p->seed_salt);
  { const char* e = getenv("GVK_PIPE_DBG"); a.dbg = e ? atoi(e) : 0; }
  a.trace = (a.dbg & 32) ? trace_buffer() : nullptr;
  const int grid = std::min(a.num_items, sm_count());
  if (p->drop_p > 0.f)
    mhsa_fwd2_kernel<true><<<grid, kThreads, kSmem, stream>>>(tqkv, a);
  else
    mhsa_fwd2_kernel<false><<<grid, kThreads, kSmem, stream>>>(tqkv, a);
  GVK_CHECK_LAUNCH("mhsa_fwd2");
  return GVK_OK;
}

}  // namespace gvk
