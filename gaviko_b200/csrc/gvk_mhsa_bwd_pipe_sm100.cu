// gvk_mhsa_bwd_pipe_sm100.cu — software-pipelined flash-attention BACKWARD for the frozen MHSA core (head dim 64, bf16 operands, fp32 softmax).
//
// Two persistent kernels, one CTA per SM, 12 warps: warp 0 TMA producer, warps 1 / 2 MMA issuers (scores / accumulators), warp 3 TMEM allocation, warps 4-11 two softmax
// groups.  A work item is ONE 128-row tile (queries in the dQ kernel, keys in the dK/dV kernel) against a stream of 128-row blocks of the
// other side, so the score MMAs run at N = 128 (64 clk for 128 x 128 x 16 — the math rate — instead of the 45-48 clk issue floor an N = 64
// instruction pays for half the work), and the two softmax groups split the 128 score COLUMNS of the block (64 each, thread = row).
//
//   dQ  kernel, per key block j:   S = Q K_j^T,  dP = dO V_j^T   ->  dS = P o (dP - delta), P = 2^(S c - lse2)   ->  dQ += dS K_j
//   dKV kernel, per query block i: S^T = K Q_i^T, dP^T = V dO_i^T ->  P^T, dS^T                                   ->  dV += P^T dO_i, dK += dS^T Q_i
//
// What makes it faster than the ping-pong form (gvk_mhsa_bwd_ws_sm100.cu) is the dependency chain: there every tile ran
// MMA(S, dP) -> softmax -> MMA(acc) -> MMA(S, dP) ... serially (two tiles hid half of it).  Here S is double-buffered in TMEM and issued two
// blocks ahead, dP is re-issued as soon as the softmax warps have pulled the previous block into registers (dQ kernel) or right behind the
// accumulating MMA that consumed it (dKV kernel), so the tensor pipe never waits for the exponentials:
//   TMEM (dQ):  S0 [0,128)  S1 [128,256)  dP [256,384)  dQ [384,448)                       dS (bf16) overwrites S in place
//   TMEM (dKV): S0 [0,128)  S1 [128,256)  dP [256,384)  dV [384,448)  dK [448,512)         P^T over S^T, dS^T over dP^T
// The 9-row tail of T = 1033 costs one N = 16 score MMA and one K = 16 accumulating MMA per item instead of a padded 64-wide tile.
// delta = rowsum(dO o O) comes from the TMA-staged dO / O tiles (both swizzled alike, and a row sum does not care about the order).
//
// No masks: TMA zero-fills rows past T, so out-of-range keys add dS * 0 to dQ and out-of-range queries carry lse2 = +inf (P = 0) in the
// statistics workspace the dQ kernel writes for the dKV kernel.  Deterministic: no atomics anywhere.
// Replaces the autograd backward of model/vision_transformer.py:65-71.
#include <algorithm>
#include <cstdlib>

#include "gvk_common.cuh"

namespace gvk {

namespace pb {
constexpr int kThreads = 384;
constexpr int kD = 64;
constexpr int kTile = 128;
constexpr int kTileBytes = kTile * kD * 2;   // 16 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kColS = 0, kColDP = 256, kColAcc = 384;

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
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
// Barrier wait of a whole warp through ONE polling lane (32 lanes spinning on the same mbarrier cost ~150 clk even when the phase is already
// complete); __syncwarp orders the other lanes' later accesses behind lane 0's acquire.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}

// Shared-memory descriptors are kept as (lo, hi) halves: hi is one constant for every SWIZZLE_128B tile with 1024-byte atoms (SBO = 1024,
// version 1), lo = (address >> 4) | LBO field, so stepping along K is ONE 32-bit add per operand — the generic make_sw128_desc costs ~7
// uniform-datapath instructions per descriptor, and at N = 64 .. 128 that set-up, not the tensor pipe, set the pace of the issuing thread.
constexpr uint32_t kDescHi = 0x40004040u;                    // SBO 1024 B | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }       // K-major tile (LBO unused = 16 B)
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (512u << 16); }    // MN-major tile (LBO 8192 B)
__device__ __forceinline__ uint64_t desc64(uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi));
  return d;
}
// D[tmem, 128 x n] = A[128 x 64 K-major smem tile] * B[n x 64 K-major smem tile]^T        (n = 16 .. 128; idesc built by the caller)
__device__ __forceinline__ void mma_scores(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, desc64(a_lo + 2 * k), desc64(b_lo + 2 * k), idesc, k > 0 ? 1u : 0u);   // 32 bytes per 16-deep K step
}
// D[tmem, 128 x 64] (+)= A * B over `ksteps` 16-deep steps: A = bf16 rows in TMEM (group g of 64 k-values sits at column 64 g of the score
// buffer it overwrote, 8 columns per step), B = [k rows x 64] smem tile with the 64 n-values contiguous (MN-major, 2048 bytes per step)
__device__ __forceinline__ void mma_accum(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, int ksteps, bool accumulate) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
  if (ksteps == 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_bf16_ts(d_tmem, a_tmem + (k < 4 ? 8 * k : 64 + 8 * (k - 4)), desc64(b_lo + 128 * k), idesc, (accumulate || k > 0) ? 1u : 0u);
  } else {
    for (int k = 0; k < ksteps; ++k) umma_bf16_ts(d_tmem, a_tmem + (k < 4 ? 8 * k : 64 + 8 * (k - 4)), desc64(b_lo + 128 * k), idesc, (accumulate || k > 0) ? 1u : 0u);
  }
}

struct Args {
  int B, T, H, dim;
  int Tpad;        // T rounded up to 128: row stride of the statistics workspace
  int nb;          // 128-row blocks per sequence
  int ntail;       // valid rows of the last block rounded up to 16
  float scale;
  const float* lse;
  float* stats;    // [2][B*H][Tpad]: delta, then the log2-domain lse (+inf past T)
  __nv_bfloat16* dqkv;
  int ld_dqkv;
  int num_items;   // B * H * nb
  MhsaDrop drop;   // attention-probability dropout of the forward call (kDrop instantiations)
  uint32_t* mask;  // [B*H][nb (key block)][Tpad (query)][4]: keep bits of the 128 keys of a block, written by the dQ kernel for the dK/dV kernel
  uint32_t* trace;
  int dbg;         // timing experiments only (GVK_PIPE_DBG): 1 = no exp2, 2 = softmax warps only pass the barriers on (results are wrong); 4 = record the timeline of CTA 0
};

// =================================================================================================
// dQ kernel
// =================================================================================================
namespace dq {
constexpr int kStages = 3;
enum { BAR_Q_FULL = 0 /*[2]*/, BAR_Q_EMPTY = 2 /*[2]*/, BAR_KV_FULL = 4, BAR_KV_EMPTY = BAR_KV_FULL + kStages, BAR_S_FULL = BAR_KV_EMPTY + kStages /*[2]*/,
       BAR_DP_FULL = BAR_S_FULL + 2, BAR_DP_FREE, BAR_DS_FULL /*[2]*/, BAR_DQ_STEP = BAR_DS_FULL + 2 /*[2]*/, BAR_DQ_FULL = BAR_DQ_STEP + 2, BAR_COUNT };
constexpr int kSmem = 6 * kTileBytes /*Q, dO, O x 2*/ + 2 * kStages * kTileBytes /*K, V ring*/ + BAR_COUNT * 8 + 64 + 1024;

template <bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
mhsa_bwd_dq_pipe_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do, const __grid_constant__ CUtensorMap tma_o, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // [2][16 KB]
  uint8_t* sdO = sQ + 2 * kTileBytes;              // [2][16 KB]
  uint8_t* sO = sdO + 2 * kTileBytes;              // [2][16 KB]
  uint8_t* sK = sO + 2 * kTileBytes;               // [kStages][16 KB]
  uint8_t* sV = sK + kStages * kTileBytes;         // [kStages][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStages * kTileBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, dim = a.dim, nb = a.nb;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    tma_prefetch_desc(&tma_o);
    for (int i = 0; i < BAR_COUNT; ++i) {
      int count = 1;
      if (i == BAR_DP_FREE || i == BAR_DS_FULL || i == BAR_DS_FULL + 1) count = 8;   // one arrival per softmax warp
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
        const int bh = item / nb, qt = item - bh * nb;
        const int h = bh % a.H, b = bh / a.H;
        const int q0 = qt * kTile, ib = work & 1;
        mbar_wait(&bars[BAR_Q_EMPTY + ib], ((work >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[BAR_Q_FULL + ib], 3 * kTileBytes);
        tma_load_3d(sQ + ib * kTileBytes, &tma_qkv, &bars[BAR_Q_FULL + ib], h * kD, q0, b);
        tma_load_3d(sdO + ib * kTileBytes, &tma_do, &bars[BAR_Q_FULL + ib], h * kD, q0, b);
        tma_load_3d(sO + ib * kTileBytes, &tma_o, &bars[BAR_Q_FULL + ib], h * kD, q0, b);
        for (int j = 0; j < nb; ++j, ++kv_iter) {
          const int st = kv_iter % kStages;
          mbar_wait(&bars[BAR_KV_EMPTY + st], ((kv_iter / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[BAR_KV_FULL + st], 2 * kTileBytes);
          tma_load_3d(sK + st * kTileBytes, &tma_qkv, &bars[BAR_KV_FULL + st], dim + h * kD, j * kTile, b);
          tma_load_3d(sV + st * kTileBytes, &tma_qkv, &bars[BAR_KV_FULL + st], 2 * dim + h * kD, j * kTile, b);
        }
      }
    } else if (warp == 1) {
      // ===================== score issuer: S(j+2), dP(j+1) (warp-uniform control flow; only the tcgen05 instructions sit under elect_one) =====
      // The instruction stream of ONE issuing thread (~25 clk per tcgen05.mma plus ~40 per commit and ~70 per barrier wait) was the bound of
      // this kernel with 16 MMAs per block; warp 1 issues the score MMAs, warp 2 the accumulating ones, ordered by barriers where the tensor
      // pipe's in-order execution used to order them.
      uint32_t kv_base = 0, work = 0, g0 = 0;   // g = g0 + j numbers the score blocks of this CTA: S buffer g & 1
      const uint32_t idesc_full = make_idesc_bf16(128, kTile, 0, 0), idesc_tail = make_idesc_bf16(128, a.ntail, 0, 0);
      Tracer tr; tr.init(a.trace, 0, (a.dbg & 4) && lane == 0);
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, kv_base += nb, g0 += nb) {
        const int ib = work & 1;
        const uint32_t q_lo = desc_lo_k(smem_u32(sQ + ib * kTileBytes)), do_lo = desc_lo_k(smem_u32(sdO + ib * kTileBytes));
        mbar_wait(&bars[BAR_Q_FULL + ib], (work >> 1) & 1);
        auto issue_s = [&](int j) {     // S(j) = Q K_j^T into buffer g & 1 once dQ(g-2), which read dS from it, is complete
          const uint32_t it = kv_base + j, g = g0 + j;
          const int st = it % kStages;
          mbar_wait(&bars[BAR_KV_FULL + st], (it / kStages) & 1);
          if (g >= 2) mbar_wait(&bars[BAR_DQ_STEP + (g & 1)], ((g >> 1) - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            mma_scores(tmem + kColS + (g & 1) * 128, q_lo, desc_lo_k(smem_u32(sK + st * kTileBytes)), j == nb - 1 ? idesc_tail : idesc_full);
            umma_commit(&bars[BAR_S_FULL + (g & 1)]);
          }
          __syncwarp();
        };
        auto issue_dp = [&](int j) {    // dP(j) = dO V_j^T once the softmax warps hold dP(j-1) in registers
          const uint32_t it = kv_base + j, g = g0 + j;
          const int st = it % kStages;
          if (g > 0) {
            mbar_wait(&bars[BAR_DP_FREE], (g - 1) & 1);
            tc_fence_after();
          }
          if (elect_one()) {
            mma_scores(tmem + kColDP, do_lo, desc_lo_k(smem_u32(sV + st * kTileBytes)), j == nb - 1 ? idesc_tail : idesc_full);
            umma_commit(&bars[BAR_DP_FULL]);
          }
          __syncwarp();
        };
        issue_s(0);
        issue_dp(0);
        if (nb > 1) issue_s(1);
        for (int j = 0; j < nb; ++j) {
          tr(0x100 + j);
          if (j + 1 < nb) issue_dp(j + 1);
          tr(0x200 + j);
          if (j + 2 < nb) issue_s(j + 2);
        }
      }
    } else if (warp == 2) {
      // ===================== accumulator issuer: dQ(j) =====================
      uint32_t kv_base = 0, work = 0, g0 = 0;
      Tracer tr; tr.init(a.trace, 3, (a.dbg & 4) && lane == 0);
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, kv_base += nb, g0 += nb) {
        const int ib = work & 1;
        for (int j = 0; j < nb; ++j) {
          const uint32_t it = kv_base + j, g = g0 + j;
          const int st = it % kStages;
          tr(0x100 + j);
          mbar_wait(&bars[BAR_DS_FULL + (g & 1)], (g >> 1) & 1);     // implies S(j), dP(j) complete, i.e. K_j / V_j landed and Q / dO of the item too
          tc_fence_after();
          tr(0x200 + j);
          if (elect_one()) {
            mma_accum(tmem + kColAcc, tmem + kColS + (g & 1) * 128, desc_lo_mn(smem_u32(sK + st * kTileBytes)), (j == nb - 1 ? a.ntail : kTile) / 16, j > 0);   // dQ (+)= dS K_j
            umma_commit(&bars[BAR_DQ_STEP + (g & 1)]);   // S buffer g & 1 may be overwritten
            umma_commit(&bars[BAR_KV_EMPTY + st]);       // K_j: S(j), dQ(j);  V_j: dP(j)
            if (j == nb - 1) {
              umma_commit(&bars[BAR_DQ_FULL]);
              umma_commit(&bars[BAR_Q_EMPTY + ib]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== softmax groups: thread = query row, group = 64 of the 128 key columns =====================
    reg_alloc<208>();
    const int grp = (warp - 4) >> 2;
    const int r = (tid - 128) & 127;          // row inside the tile = TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tdP = tmem + kColDP + lane_off + 64 * grp, tdQ = tmem + kColAcc + lane_off + 32 * grp;
    const float c2 = a.scale * kLog2e;
    const float2 c2v = make_float2(c2, c2);
    uint32_t g = 0, work = 0;
    const MhsaDrop drop = kDrop ? mhsa_salted(a.drop) : a.drop;
    Tracer tr; tr.init(a.trace, 1 + grp, (a.dbg & 4) && (warp & 3) == 0 && lane == 0);
    // Per-item set-up (row statistics), software-pipelined one item ahead: the set-up of item n+1 runs in the shadow of item n's last dQ MMAs
    // (the groups would otherwise idle there), so block 0 of the next item starts right behind the epilogue.
    struct ItemCtx { int b, h, bh, row; float lse2, delta; };
    float lse_pre = 0.f;      // lse of the item AFTER the one being set up: its global load has a whole item to land
    auto lse_of = [&](int item) {
      const int bh = item / nb, row = (item - bh * nb) * kTile + r;
      return (item < a.num_items && row < T) ? a.lse[(size_t)bh * T + row] : 0.f;
    };
    auto setup = [&](int item, uint32_t wk) {
      ItemCtx c;
      const int bh = item / nb, qt = item - bh * nb;
      c.h = bh % a.H;
      c.b = bh / a.H;
      c.bh = bh;
      c.row = qt * kTile + r;
      const int ib = wk & 1;
      c.lse2 = c.row < T ? lse_pre * kLog2e : INFINITY;
      lse_pre = lse_of(item + gridDim.x);
      mbar_wait_warp(&bars[BAR_Q_FULL + ib], (wk >> 1) & 1, lane);
      const uint4* pd = reinterpret_cast<const uint4*>(sdO + ib * kTileBytes + r * 128);
      const uint4* po = reinterpret_cast<const uint4*>(sO + ib * kTileBytes + r * 128);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int cch = 0; cch < 8; ++cch) {
        const int cc = (cch + r) & 7;             // rotate the 16-byte chunk with the row: conflict-free
        const uint4 x = po[cc], y = pd[cc];
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[u]));
          const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[u]));
          d0 = fmaf(fx.x, fy.x, d0);
          d1 = fmaf(fx.y, fy.y, d1);
        }
      }
      c.delta = d0 + d1;
      if (grp == 0) {
        a.stats[(size_t)bh * a.Tpad + c.row] = c.delta;
        a.stats[(size_t)a.B * a.H * a.Tpad + (size_t)bh * a.Tpad + c.row] = c.lse2;
      }
      return c;
    };
    ItemCtx cur{};
    lse_pre = lse_of(blockIdx.x);
    if ((int)blockIdx.x < a.num_items) cur = setup(blockIdx.x, 0);
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work) {
      const int b = cur.b, h = cur.h, bh = cur.bh, row = cur.row;
      const float2 nl = make_float2(-cur.lse2, -cur.lse2), nd = make_float2(-cur.delta, -cur.delta);
      tr(0x900);
      for (int j = 0; j < nb; ++j, ++g) {
        const int ncols = (j == nb - 1 ? a.ntail : kTile) - 64 * grp;      // columns of this group in block j (<= 0: none)
        const uint32_t tS = tmem + kColS + (g & 1) * 128 + lane_off + 64 * grp;
        tr(0x100 + j);
        mbar_wait_warp(&bars[BAR_S_FULL + (g & 1)], (g >> 1) & 1, lane);
        mbar_wait_warp(&bars[BAR_DP_FULL], g & 1, lane);
        tc_fence_after();
        tr(0x300 + j);
        if (a.dbg & 2) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&bars[BAR_DP_FREE]); mbar_arrive(&bars[BAR_DS_FULL + (g & 1)]); }
          continue;
        }
        // 32 scores at a time; the second half's TMEM loads travel while the first half is in its exp2 phase
        auto chunk = [&](const float (&s)[32], const float (&dp)[32], int c) {
          uint32_t pk[16];
          uint32_t keep = 0xFFFFFFFFu;
          if (kDrop) {     // replay the forward's dropout decisions of these 32 keys and hand them to the dK/dV kernel
            keep = mhsa_keep16(drop, bh, row, 8 * j + 4 * grp + 2 * c) | (mhsa_keep16(drop, bh, row, 8 * j + 4 * grp + 2 * c + 1) << 16);
            a.mask[(((size_t)bh * nb + j) * a.Tpad + row) * 4 + 2 * grp + c] = keep;
          }
          const float2 ik = make_float2(a.drop.inv_keep, a.drop.inv_keep);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float2 e = ffma2(make_float2(s[i], s[i + 1]), c2v, nl);
            if (!(a.dbg & 1)) {
              e.x = fast_ex2(e.x);
              e.y = fast_ex2(e.y);
            }
            float2 dpe = make_float2(dp[i], dp[i + 1]);
            if (kDrop) {   // d(P dropped) / dP = keep / (1 - p)
              dpe = fmul2(dpe, ik);
              if (!((keep >> i) & 1u)) dpe.x = 0.f;
              if (!((keep >> (i + 1)) & 1u)) dpe.y = 0.f;
            }
            const float2 ds = fmul2(e, fadd2(dpe, nd));   // dS = P (dP - delta)
            pk[i >> 1] = pack_bf16x2(ds.x, ds.y);
          }
          tmem_st_32x16(tS + 16 * c, pk);     // packed dS over S columns this thread has already read
        };
        if (ncols > 0) {
          float s0[32], dp0[32], s1[32], dp1[32];
          tmem_ld_32x32(tdP, dp0);
          tmem_ld_32x32(tS, s0);
          tc_wait_ld();
          if (ncols > 32) {
            tmem_ld_32x32(tdP + 32, dp1);
            tmem_ld_32x32(tS + 32, s1);
          }
          chunk(s0, dp0, 0);
          if (ncols > 32) tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_DP_FREE]);     // dP(j) is in registers: dP(j+1) may overwrite it
          tr(0x400 + j);
          if (ncols > 32) chunk(s1, dp1, 1);
          tc_wait_st();
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_DP_FREE]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_DS_FULL + (g & 1)]);
        tr(0x500 + j);
      }
      if (item + (int)gridDim.x < a.num_items) cur = setup(item + gridDim.x, work + 1);
      tr(0x800);
      // ---- epilogue: dQ * scale -> bf16 (each group stores 32 of the 64 columns)
      mbar_wait_warp(&bars[BAR_DQ_FULL], work & 1, lane);
      tc_fence_after();
      tr(0x600);
      {
        float v[32];
        tmem_ld_32x32(tdQ, v);
        tc_wait_ld();
        if (row < T) {
          uint4* dst = reinterpret_cast<uint4*>(a.dqkv + ((size_t)b * T + row) * a.ld_dqkv + h * kD + 32 * grp);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 pk;
            pk.x = pack_bf16x2(v[8 * u + 0] * a.scale, v[8 * u + 1] * a.scale);
            pk.y = pack_bf16x2(v[8 * u + 2] * a.scale, v[8 * u + 3] * a.scale);
            pk.z = pack_bf16x2(v[8 * u + 4] * a.scale, v[8 * u + 5] * a.scale);
            pk.w = pack_bf16x2(v[8 * u + 6] * a.scale, v[8 * u + 7] * a.scale);
            dst[u] = pk;
          }
        }
      }
      // the next item's dQ(0) (accumulate = 0) is gated by the groups' next DS_FULL arrivals, i.e. after these reads: no extra barrier
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}
}  // namespace dq

// =================================================================================================
// dK / dV kernel
// =================================================================================================
namespace dkv {
constexpr int kStages = 4;
constexpr int kStatBytes = 2 * kTile * 4;                         // delta[128], lse2[128] of the streamed query block
constexpr int kMaskBytes = kTile * 16;                            // dropout keep bits: [128 queries][4 words = 128 keys]
constexpr int kStageBytes = 2 * kTileBytes + kStatBytes;
constexpr int kColDV = 384, kColDK = 448;
enum { BAR_KV_FULL = 0 /*[2]*/, BAR_KV_EMPTY = 2 /*[2]*/, BAR_Q_FULL = 4, BAR_Q_EMPTY = BAR_Q_FULL + kStages, BAR_S_FULL = BAR_Q_EMPTY + kStages /*[2]*/,
       BAR_DP_FULL = BAR_S_FULL + 2, BAR_P_FULL /*[2]*/, BAR_DS_FULL = BAR_P_FULL + 2 /*[2]*/, BAR_DV_STEP = BAR_DS_FULL + 2 /*[2]*/, BAR_DK_STEP = BAR_DV_STEP + 2,
       BAR_ACC_FULL, BAR_COUNT };
constexpr int kSmem = 4 * kTileBytes /*K, V x 2*/ + kStages * kStageBytes + BAR_COUNT * 8 + 64 + 1024;
constexpr int kSmemDrop = kSmem + kStages * kMaskBytes;

template <bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
mhsa_bwd_dkv_pipe_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;                              // [2][16 KB]
  uint8_t* sV = sK + 2 * kTileBytes;               // [2][16 KB]
  uint8_t* sQ = sV + 2 * kTileBytes;               // [kStages][16 KB]
  uint8_t* sdO = sQ + kStages * kTileBytes;        // [kStages][16 KB]
  float* sStat = reinterpret_cast<float*>(sdO + kStages * kTileBytes);   // [kStages][delta 128 | lse2 128]
  uint32_t* sMask = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(sStat) + kStages * kStatBytes);   // [kStages][128][4] (kDrop only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sMask) + (kDrop ? kStages * kMaskBytes : 0));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, dim = a.dim, nb = a.nb;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    for (int i = 0; i < BAR_COUNT; ++i) {
      int count = 1;
      if (i >= BAR_P_FULL && i < BAR_DV_STEP) count = 8;
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
      uint32_t q_iter = 0, work = 0;
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work) {
        const int bh = item / nb, kt = item - bh * nb;
        const int h = bh % a.H, b = bh / a.H;
        const int k0 = kt * kTile, ib = work & 1;
        mbar_wait(&bars[BAR_KV_EMPTY + ib], ((work >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[BAR_KV_FULL + ib], 2 * kTileBytes);
        tma_load_3d(sK + ib * kTileBytes, &tma_qkv, &bars[BAR_KV_FULL + ib], dim + h * kD, k0, b);
        tma_load_3d(sV + ib * kTileBytes, &tma_qkv, &bars[BAR_KV_FULL + ib], 2 * dim + h * kD, k0, b);
        const float* g_delta = a.stats + (size_t)bh * a.Tpad;
        const float* g_lse2 = a.stats + (size_t)a.B * a.H * a.Tpad + (size_t)bh * a.Tpad;
        for (int i = 0; i < nb; ++i, ++q_iter) {
          const int st = q_iter % kStages;
          mbar_wait(&bars[BAR_Q_EMPTY + st], ((q_iter / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[BAR_Q_FULL + st], kStageBytes + (kDrop ? kMaskBytes : 0));
          if (kDrop) bulk_load(sMask + st * kTile * 4, a.mask + (((size_t)bh * nb + kt) * a.Tpad + (size_t)i * kTile) * 4, kMaskBytes, &bars[BAR_Q_FULL + st]);
          tma_load_3d(sQ + st * kTileBytes, &tma_qkv, &bars[BAR_Q_FULL + st], h * kD, i * kTile, b);
          tma_load_3d(sdO + st * kTileBytes, &tma_do, &bars[BAR_Q_FULL + st], h * kD, i * kTile, b);
          bulk_load(sStat + st * 2 * kTile, g_delta + i * kTile, kTile * 4, &bars[BAR_Q_FULL + st]);
          bulk_load(sStat + st * 2 * kTile + kTile, g_lse2 + i * kTile, kTile * 4, &bars[BAR_Q_FULL + st]);
        }
      }
    } else if (warp == 1) {
      // ===================== score issuer: S^T(i+2), dP^T(i+1) =====================
      uint32_t q_base = 0, work = 0, g0 = 0;
      const uint32_t idesc_full = make_idesc_bf16(128, kTile, 0, 0), idesc_tail = make_idesc_bf16(128, a.ntail, 0, 0);
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, q_base += nb, g0 += nb) {
        const int ib = work & 1;
        const uint32_t k_lo = desc_lo_k(smem_u32(sK + ib * kTileBytes)), v_lo = desc_lo_k(smem_u32(sV + ib * kTileBytes));
        mbar_wait(&bars[BAR_KV_FULL + ib], (work >> 1) & 1);
        auto issue_s = [&](int i) {     // S^T(i) = K Q_i^T into buffer g & 1 once dV(g-2), which read P^T from it, is complete
          const uint32_t it = q_base + i, g = g0 + i;
          const int st = it % kStages;
          mbar_wait(&bars[BAR_Q_FULL + st], (it / kStages) & 1);
          if (g >= 2) mbar_wait(&bars[BAR_DV_STEP + (g & 1)], ((g >> 1) - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            mma_scores(tmem + kColS + (g & 1) * 128, k_lo, desc_lo_k(smem_u32(sQ + st * kTileBytes)), i == nb - 1 ? idesc_tail : idesc_full);
            umma_commit(&bars[BAR_S_FULL + (g & 1)]);
          }
          __syncwarp();
        };
        auto issue_dp = [&](int i) {    // dP^T(i) = V dO_i^T once dK(g-1), which read dS^T from the same columns, is complete
          const uint32_t g = g0 + i;
          const int st = (q_base + i) % kStages;
          if (g > 0) {
            mbar_wait(&bars[BAR_DK_STEP], (g - 1) & 1);
            tc_fence_after();
          }
          if (elect_one()) {
            mma_scores(tmem + kColDP, v_lo, desc_lo_k(smem_u32(sdO + st * kTileBytes)), i == nb - 1 ? idesc_tail : idesc_full);
            umma_commit(&bars[BAR_DP_FULL]);
          }
          __syncwarp();
        };
        issue_s(0);
        issue_dp(0);
        if (nb > 1) issue_s(1);
        for (int i = 0; i < nb; ++i) {
          if (i + 2 < nb) issue_s(i + 2);
          if (i + 1 < nb) issue_dp(i + 1);
        }
      }
    } else if (warp == 2) {
      // ===================== accumulator issuer: dV(i), dK(i) =====================
      uint32_t q_base = 0, work = 0, g0 = 0;
      Tracer tr; tr.init(a.trace, 3, (a.dbg & 16) && lane == 0);
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, q_base += nb, g0 += nb) {
        const int ib = work & 1;
        for (int i = 0; i < nb; ++i) {
          const uint32_t it = q_base + i, g = g0 + i;
          const int st = it % kStages;
          const int ksteps = (i == nb - 1 ? a.ntail : kTile) / 16;
          tr(0x100 + i);
          mbar_wait(&bars[BAR_P_FULL + (g & 1)], (g >> 1) & 1);       // implies S^T(i) complete: Q_i landed, and K / V of the item
          tc_fence_after();
          tr(0x200 + i);
          if (elect_one()) {
            mma_accum(tmem + kColDV, tmem + kColS + (g & 1) * 128, desc_lo_mn(smem_u32(sdO + st * kTileBytes)), ksteps, i > 0);   // dV (+)= P^T dO_i
            umma_commit(&bars[BAR_DV_STEP + (g & 1)]);
          }
          __syncwarp();
          tr(0x300 + i);
          mbar_wait(&bars[BAR_DS_FULL + (g & 1)], (g >> 1) & 1);
          tc_fence_after();
          tr(0x400 + i);
          if (elect_one()) {
            mma_accum(tmem + kColDK, tmem + kColDP, desc_lo_mn(smem_u32(sQ + st * kTileBytes)), ksteps, i > 0);                    // dK (+)= dS^T Q_i
            umma_commit(&bars[BAR_DK_STEP]);
            umma_commit(&bars[BAR_Q_EMPTY + st]);
            if (i == nb - 1) {
              umma_commit(&bars[BAR_ACC_FULL]);
              umma_commit(&bars[BAR_KV_EMPTY + ib]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== softmax groups: thread = key row, group = 64 of the 128 query columns =====================
    reg_alloc<208>();
    const int grp = (warp - 4) >> 2;
    const int r = (tid - 128) & 127;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tdP = tmem + kColDP + lane_off + 64 * grp;
    const uint32_t tAcc = tmem + (grp == 0 ? kColDK : kColDV) + lane_off;
    const float c2 = a.scale * kLog2e;
    const float2 c2v = make_float2(c2, c2);
    uint32_t g = 0, work = 0, q_base = 0;
    Tracer tr; tr.init(a.trace, 1 + grp, (a.dbg & 16) && (warp & 3) == 0 && lane == 0);
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, q_base += nb) {
      const int bh = item / nb, kt = item - bh * nb;
      const int h = bh % a.H, b = bh / a.H;
      for (int i = 0; i < nb; ++i, ++g) {
        tr(0x100 + i);
        const uint32_t it = q_base + i;
        const int st = it % kStages;
        const int ncols = (i == nb - 1 ? a.ntail : kTile) - 64 * grp;
        const float* s_delta = sStat + st * 2 * kTile + 64 * grp;
        const float* s_lse2 = s_delta + kTile;
        const uint32_t tS = tmem + kColS + (g & 1) * 128 + lane_off + 64 * grp;
        mbar_wait_warp(&bars[BAR_Q_FULL + st], (it / kStages) & 1, lane);       // the statistics of the stage (bulk copies) are visible
        mbar_wait_warp(&bars[BAR_S_FULL + (g & 1)], (g >> 1) & 1, lane);
        tc_fence_after();
        tr(0x200 + i);
        if (a.dbg & 2) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_P_FULL + (g & 1)]);
          mbar_wait_warp(&bars[BAR_DP_FULL], g & 1, lane);
          if (lane == 0) mbar_arrive(&bars[BAR_DS_FULL + (g & 1)]);
          continue;
        }
        float p0[32], p1[32];
        const uint32_t* s_mask = sMask + (st * kTile + 64 * grp) * 4 + (warp & 3);     // this thread's key is bit `lane` of word (warp & 3) of each query
        auto kept = [&](int q) { return kDrop ? ((s_mask[4 * q] >> lane) & 1u) != 0u : true; };
        auto p_chunk = [&](float (&p)[32], int c) {      // P^T = 2^(S^T c - lse2[q]) in place, packed copy over S^T columns already read
          uint32_t pk[16];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 l4 = *reinterpret_cast<const float4*>(s_lse2 + 32 * c + 4 * q4);
            const int q = 4 * q4;
            float2 e0 = ffma2(make_float2(p[q], p[q + 1]), c2v, make_float2(-l4.x, -l4.y));
            float2 e1 = ffma2(make_float2(p[q + 2], p[q + 3]), c2v, make_float2(-l4.z, -l4.w));
            if (!(a.dbg & 1)) { e0.x = fast_ex2(e0.x); e0.y = fast_ex2(e0.y); e1.x = fast_ex2(e1.x); e1.y = fast_ex2(e1.y); }
            p[q] = e0.x;
            p[q + 1] = e0.y;
            p[q + 2] = e1.x;
            p[q + 3] = e1.y;
            if (kDrop) {      // dV sees the dropped probabilities (their 1 / (1 - p) is applied in the epilogue)
              if (!kept(32 * c + q)) e0.x = 0.f;
              if (!kept(32 * c + q + 1)) e0.y = 0.f;
              if (!kept(32 * c + q + 2)) e1.x = 0.f;
              if (!kept(32 * c + q + 3)) e1.y = 0.f;
            }
            pk[2 * q4] = pack_bf16x2(e0.x, e0.y);
            pk[2 * q4 + 1] = pack_bf16x2(e1.x, e1.y);
          }
          tmem_st_32x16(tS + 16 * c, pk);
        };
        auto ds_chunk = [&](const float (&p)[32], const float (&dp)[32], int c) {   // dS^T = P^T (dP^T - delta[q]) over dP^T columns already read
          uint32_t pk[16];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 d4 = *reinterpret_cast<const float4*>(s_delta + 32 * c + 4 * q4);
            const int q = 4 * q4;
            float2 e0 = make_float2(dp[q], dp[q + 1]), e1 = make_float2(dp[q + 2], dp[q + 3]);
            if (kDrop) {
              const float2 ik = make_float2(a.drop.inv_keep, a.drop.inv_keep);
              e0 = fmul2(e0, ik);
              e1 = fmul2(e1, ik);
              if (!kept(32 * c + q)) e0.x = 0.f;
              if (!kept(32 * c + q + 1)) e0.y = 0.f;
              if (!kept(32 * c + q + 2)) e1.x = 0.f;
              if (!kept(32 * c + q + 3)) e1.y = 0.f;
            }
            const float2 a0 = fmul2(make_float2(p[q], p[q + 1]), fadd2(e0, make_float2(-d4.x, -d4.y)));
            const float2 a1 = fmul2(make_float2(p[q + 2], p[q + 3]), fadd2(e1, make_float2(-d4.z, -d4.w)));
            pk[2 * q4] = pack_bf16x2(a0.x, a0.y);
            pk[2 * q4 + 1] = pack_bf16x2(a1.x, a1.y);
          }
          tmem_st_32x16(tdP + 16 * c, pk);
        };
        if (ncols > 0) {
          tmem_ld_32x32(tS, p0);
          tc_wait_ld();
          if (ncols > 32) tmem_ld_32x32(tS + 32, p1);
          p_chunk(p0, 0);
          if (ncols > 32) {
            tc_wait_ld();
            p_chunk(p1, 1);
          }
          tc_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_P_FULL + (g & 1)]);
        tr(0x300 + i);
        mbar_wait_warp(&bars[BAR_DP_FULL], g & 1, lane);
        tc_fence_after();
        tr(0x400 + i);
        if (ncols > 0) {
          float dp0[32], dp1[32];
          tmem_ld_32x32(tdP, dp0);
          tc_wait_ld();
          if (ncols > 32) tmem_ld_32x32(tdP + 32, dp1);
          ds_chunk(p0, dp0, 0);
          if (ncols > 32) {
            tc_wait_ld();
            ds_chunk(p1, dp1, 1);
          }
          tc_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_DS_FULL + (g & 1)]);
        tr(0x500 + i);
      }
      tr(0x600);
      // ---- epilogue: group 0 stores dK * scale, group 1 dV
      mbar_wait_warp(&bars[BAR_ACC_FULL], work & 1, lane);
      tc_fence_after();
      const int row = kt * kTile + r;
      const float sc = grp == 0 ? a.scale : (kDrop ? a.drop.inv_keep : 1.0f);
      __nv_bfloat16* base = a.dqkv + ((size_t)b * T + row) * a.ld_dqkv + (grp == 0 ? dim : 2 * dim) + h * kD;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld_32x32(tAcc + 32 * c, v);
        tc_wait_ld();
        if (row < T) {
          uint4* dst = reinterpret_cast<uint4*>(base + 32 * c);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 pk;
            pk.x = pack_bf16x2(v[8 * u + 0] * sc, v[8 * u + 1] * sc);
            pk.y = pack_bf16x2(v[8 * u + 2] * sc, v[8 * u + 3] * sc);
            pk.z = pack_bf16x2(v[8 * u + 4] * sc, v[8 * u + 5] * sc);
            pk.w = pack_bf16x2(v[8 * u + 6] * sc, v[8 * u + 7] * sc);
            dst[u] = pk;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}
}  // namespace dkv
}  // namespace pb

size_t mhsa_bwd_mask_words(int B, int T, int H) {
  const size_t nb = (size_t)(T + pb::kTile - 1) / pb::kTile;
  return (size_t)B * H * nb * (nb * pb::kTile) * 4;
}

static uint32_t* g_trace_dev = nullptr;
uint32_t* trace_buffer() {
  if (!g_trace_dev) {
    if (cudaMalloc(&g_trace_dev, sizeof(uint32_t) * kTraceRoles * kTraceN * 2) != cudaSuccess) return nullptr;
    cudaMemset(g_trace_dev, 0, sizeof(uint32_t) * kTraceRoles * kTraceN * 2);
  }
  return g_trace_dev;
}
int debug_trace(uint32_t* out, int n_words) {
  if (!g_trace_dev) { set_last_error("gvk_debug_trace: no trace was recorded (GVK_PIPE_DBG)"); return GVK_ERR_INVALID_ARGUMENT; }
  const size_t bytes = std::min<size_t>(sizeof(uint32_t) * kTraceRoles * kTraceN * 2, (size_t)n_words * 4);
  return cuda_status(cudaMemcpy(out, g_trace_dev, bytes, cudaMemcpyDeviceToHost), "debug_trace");
}

int mhsa_bwd_pipe(const gvk_mhsa_bwd_params* p, cudaStream_t stream) {
  using namespace pb;
  static bool configured = false;
  if (!configured) {
    int st = cuda_status(cudaFuncSetAttribute(dq::mhsa_bwd_dq_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq::kSmem), "mhsa_bwd_dq_pipe smem");
    if (st != GVK_OK) return st;
    st = cuda_status(cudaFuncSetAttribute(dq::mhsa_bwd_dq_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq::kSmem), "mhsa_bwd_dq_pipe smem");
    if (st != GVK_OK) return st;
    st = cuda_status(cudaFuncSetAttribute(dkv::mhsa_bwd_dkv_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv::kSmem), "mhsa_bwd_dkv_pipe smem");
    if (st != GVK_OK) return st;
    st = cuda_status(cudaFuncSetAttribute(dkv::mhsa_bwd_dkv_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv::kSmemDrop), "mhsa_bwd_dkv_pipe smem");
    if (st != GVK_OK) return st;
    configured = true;
  }
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(p->out) & 15) == 0, "gvk_mhsa_bwd: out must be 16-byte aligned");
  const int dim = p->H * kD;
  CUtensorMap tqkv, tdo, to;
  int st = make_tma_3d_bf16(&tqkv, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, kTile, kD);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tdo, p->dout, p->B, p->T, dim, p->ld_dout, (uint64_t)p->T * p->ld_dout, kTile, kD);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&to, p->out, p->B, p->T, dim, p->ld_out, (uint64_t)p->T * p->ld_out, kTile, kD);
  if (st != GVK_OK) return st;
  Args a;
  a.B = p->B; a.T = p->T; a.H = p->H; a.dim = dim; a.scale = p->scale;
  a.nb = (p->T + kTile - 1) / kTile;
  a.Tpad = a.nb * kTile;
  a.ntail = (p->T - (a.nb - 1) * kTile + 15) / 16 * 16;
  a.lse = p->lse;
  a.stats = p->delta;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(p->dqkv);
  a.ld_dqkv = p->ld_dqkv;
  a.num_items = p->B * p->H * a.nb;
  { const char* e = getenv("GVK_PIPE_DBG"); a.dbg = e ? atoi(e) : 0; }
  a.trace = (a.dbg & (4 | 16)) ? trace_buffer() : nullptr;
  const int grid = std::min(a.num_items, sm_count());
  const bool drop = p->drop_p > 0.f;
  GVK_CHECK_ARG(p->drop_p >= 0.f && p->drop_p < 1.f, "gvk_mhsa_bwd: drop_p must be in [0, 1)");
  GVK_CHECK_ARG(!drop || (p->mask_ws && (reinterpret_cast<uintptr_t>(p->mask_ws) & 15) == 0), "gvk_mhsa_bwd: dropout needs a 16-byte aligned mask workspace");
  a.drop = make_mhsa_drop(p->drop_p, p->seed, p->seed_salt);
  a.mask = p->mask_ws;
  if (drop) {
    dq::mhsa_bwd_dq_pipe_kernel<true><<<grid, kThreads, dq::kSmem, stream>>>(tqkv, tdo, to, a);
    GVK_CHECK_LAUNCH("mhsa_bwd_dq_pipe");
    dkv::mhsa_bwd_dkv_pipe_kernel<true><<<grid, kThreads, dkv::kSmemDrop, stream>>>(tqkv, tdo, a);
    GVK_CHECK_LAUNCH("mhsa_bwd_dkv_pipe");
  } else {
    dq::mhsa_bwd_dq_pipe_kernel<false><<<grid, kThreads, dq::kSmem, stream>>>(tqkv, tdo, to, a);
    GVK_CHECK_LAUNCH("mhsa_bwd_dq_pipe");
    dkv::mhsa_bwd_dkv_pipe_kernel<false><<<grid, kThreads, dkv::kSmem, stream>>>(tqkv, tdo, a);
    GVK_CHECK_LAUNCH("mhsa_bwd_dkv_pipe");
  }
  return GVK_OK;
}

}  // namespace gvk
