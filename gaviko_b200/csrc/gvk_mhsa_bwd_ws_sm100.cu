// gvk_mhsa_bwd_ws_sm100.cu — warp-specialised flash-attention BACKWARD for the frozen MHSA core (head dim 64, bf16 operands, fp32 softmax).
//
// Two persistent kernels with the forward kernel's structure (gvk_mhsa_ws_sm100.cu), one CTA per SM, 12 warps:
//   warp 0      TMA producer
//   warps 1, 2  MMA issuers (one per tile of the pair; only the tcgen05 instructions sit under elect_one)
//   warp 3      TMEM allocation
//   warps 4-7   softmax group A   (thread t owns row t of tile A = TMEM lane t)
//   warps 8-11  softmax group B   (tile B; the two tiles ping-pong: the tensor pipe works on one tile's MMAs while the other tile's
//                                  group is in its exp2 phase, so neither the MMA issue chain nor the SFU waits for the other)
//
//   dQ  kernel: a work item = (volume, head, pair of 128-row query tiles); loop over 64-key K/V tiles:
//                 S = Q K^T, dP = dO V^T  ->  dS = P o (dP - delta), P = 2^(S c - lse2)  ->  dQ += dS K          (3 MMA groups / step)
//               also writes delta = rowsum(dO o O) and the log2-domain lse into the (padded) workspace for the second kernel.
//   dKV kernel: a work item = (volume, head, pair of 128-row key tiles); loop over 64-row query tiles:
//                 S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T  ->  dV += P^T dO, dK += dS^T Q                      (4 MMA groups / step)
// Two kernels instead of one: dQ needs no atomics and every result is deterministic (5 MMA groups would become 7, but the dQ reduction
// across key tiles would need fp32 atomics on a [B*T, dim] buffer plus a conversion pass).
//
// No masks are needed: TMA zero-fills rows past T, so out-of-range keys contribute dS * 0 to dQ and out-of-range queries have lse2 = +inf
// (P = 0) — written that way into the workspace by the dQ kernel, whose row range covers the padded length of the dKV kernel's query tiles.
// P / dS go registers -> TMEM (packed bf16 over the columns S / dP occupied) and feed the accumulating MMAs as their A operand.
//
// Replaces the autograd backward of model/vision_transformer.py:65-71 (which keeps two (B, H, T, T) matrices alive per layer).
#include <algorithm>
#include <cstdlib>

#include "gvk_common.cuh"

namespace gvk {

namespace wsb {
constexpr int kThreads = 384;
constexpr int kD = 64;
constexpr int kTile = 128;                   // rows of the resident tiles (queries in the dQ kernel, keys in the dKV kernel)
constexpr int kStep = 64;                    // rows of the streamed tiles
constexpr int kTileBytes = kTile * kD * 2;   // 16 KB
constexpr int kStepBytes = kStep * kD * 2;   // 8 KB
constexpr float kLog2e = 1.4426950408889634f;

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
// 1-D bulk copy global -> shared with mbarrier completion (16-byte aligned addresses, size a multiple of 16)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
// D[tmem, 128 x 64] (+)= A[128 x 64 K-major smem tile] * B[64 x 64 K-major smem tile]^T
__device__ __forceinline__ void mma_kk(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, bool accumulate) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16(d_tmem, make_sw128_desc(a_addr + k * 32, 16, 1024), make_sw128_desc(b_addr + k * 32, 16, 1024), idesc, (accumulate || k > 0) ? 1u : 0u);
}
// D[tmem, 128 x 64] (+)= A * B, A = [128 x 64] bf16 in TMEM (32 columns), B = [64 (k) rows x 64 (n)] smem tile, n contiguous (MN-major)
__device__ __forceinline__ void mma_tmn(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, bool accumulate) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16_ts(d_tmem, a_tmem + 8 * k, make_sw128_desc(b_addr + k * 2048, 8192, 1024), idesc, (accumulate || k > 0) ? 1u : 0u);
}

struct Args {
  int B, T, H, dim;
  int Tpad;        // T rounded up to the streamed tile (64): row stride of the statistics workspace
  float scale;
  const __nv_bfloat16* out;
  int ld_out;
  const __nv_bfloat16* dout;
  int ld_dout;
  const float* lse;
  float* stats;    // [2][B*H][Tpad]: delta, then the log2-domain lse (+inf past T)
  __nv_bfloat16* dqkv;
  int ld_dqkv;
  int pairs;       // tile pairs per (volume, head)
  int num_items;   // B * H * pairs
};

// =================================================================================================
// dQ kernel
// =================================================================================================
namespace dq {
constexpr int kStages = 6;
// TMEM columns per tile X: S at X*192 (dS overwrites its first 32), dP at X*192 + 64, dQ at X*192 + 128
constexpr int kTileCols = 192;
enum { BAR_Q_FULL = 0, BAR_Q_EMPTY = 1, BAR_KV_FULL = 2, BAR_KV_EMPTY = BAR_KV_FULL + kStages, BAR_S_FULL = BAR_KV_EMPTY + kStages /*[X]*/,
       BAR_DS_FULL = BAR_S_FULL + 2 /*[X]*/, BAR_DQ_FULL = BAR_DS_FULL + 2 /*[X]*/, BAR_COUNT = BAR_DQ_FULL + 2 };
constexpr int kSmem = 4 * kTileBytes /*Q A,B, dO A,B*/ + 2 * kStages * kStepBytes /*K,V ring*/ + BAR_COUNT * 8 + 64 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
mhsa_bwd_dq_ws_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv, const __grid_constant__ CUtensorMap tma_do, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // [2][16 KB]
  uint8_t* sdO = sQ + 2 * kTileBytes;              // [2][16 KB]
  uint8_t* sK = sdO + 2 * kTileBytes;              // [kStages][8 KB]
  uint8_t* sV = sK + kStages * kStepBytes;         // [kStages][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStages * kStepBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, dim = a.dim;
  const int nkv = (T + kStep - 1) / kStep;

  if (tid == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    tma_prefetch_desc(&tma_do);
    for (int i = 0; i < BAR_COUNT; ++i) {
      int count = 1;
      if (i >= BAR_DS_FULL && i < BAR_DS_FULL + 2) count = 4;                                   // one arrival per softmax warp
      if (i == BAR_Q_EMPTY || (i >= BAR_KV_EMPTY && i < BAR_KV_EMPTY + kStages)) count = 2;      // one per MMA issuer
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
        mbar_arrive_expect_tx(&bars[BAR_Q_FULL], activeB ? 4 * kTileBytes : 2 * kTileBytes);
        tma_load_3d(sQ, &tma_q, &bars[BAR_Q_FULL], h * kD, q0, b);
        tma_load_3d(sdO, &tma_do, &bars[BAR_Q_FULL], h * kD, q0, b);
        if (activeB) {
          tma_load_3d(sQ + kTileBytes, &tma_q, &bars[BAR_Q_FULL], h * kD, q0 + kTile, b);
          tma_load_3d(sdO + kTileBytes, &tma_do, &bars[BAR_Q_FULL], h * kD, q0 + kTile, b);
        }
        for (int j = 0; j < nkv; ++j, ++kv_iter) {
          const int st = kv_iter % kStages;
          mbar_wait(&bars[BAR_KV_EMPTY + st], ((kv_iter / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[BAR_KV_FULL + st], 2 * kStepBytes);
          tma_load_3d(sK + st * kStepBytes, &tma_kv, &bars[BAR_KV_FULL + st], dim + h * kD, j * kStep, b);
          tma_load_3d(sV + st * kStepBytes, &tma_kv, &bars[BAR_KV_FULL + st], 2 * dim + h * kD, j * kStep, b);
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: warp 1 drives tile A, warp 2 tile B =====================
      const int X = warp - 1;
      const uint32_t q_addr = smem_u32(sQ + X * kTileBytes), do_addr = smem_u32(sdO + X * kTileBytes);
      const uint32_t tS = tmem + X * kTileCols, tdP = tS + 64, tdQ = tS + 128;
      uint32_t kv_base = 0, work = 0;
      uint32_t steps = 0;   // softmax steps of this tile so far (parity of S_FULL / DS_FULL)
      auto issue_s = [&](uint32_t it) {   // S = Q K(it)^T and dP = dO V(it)^T, 128 x 64 x 64 each
        const int st = it % kStages;
        mbar_wait(&bars[BAR_KV_FULL + st], (it / kStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          mma_kk(tS, q_addr, smem_u32(sK + st * kStepBytes), false);
          mma_kk(tdP, do_addr, smem_u32(sV + st * kStepBytes), false);
          umma_commit(&bars[BAR_S_FULL + X]);
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, kv_base += nkv) {
        const int qp = item % a.pairs;
        const bool activeB = qp * 2 * kTile + kTile < T;
        const int releases = (X == 0 && !activeB) ? 2 : 1;
        // both issuers observe EVERY Q_FULL phase, even for an item tile B sits out (a skipped phase would alias with the one before it)
        mbar_wait(&bars[BAR_Q_FULL], work & 1);
        tc_fence_after();
        if (X == 1 && !activeB) continue;       // warp 1 then releases the shared stages for both
        issue_s(kv_base);
        for (int j = 0; j < nkv; ++j, ++steps) {
          const uint32_t itj = kv_base + j;
          const int st = itj % kStages;
          mbar_wait(&bars[BAR_DS_FULL + X], steps & 1);
          tc_fence_after();
          if (elect_one()) mma_tmn(tdQ, tS, smem_u32(sK + st * kStepBytes), j > 0);     // dQ (+)= dS K(j); reads dS before S(j+1) overwrites it (in order)
          __syncwarp();
          if (j + 1 < nkv) {
            issue_s(itj + 1);
          } else if (elect_one()) {
            umma_commit(&bars[BAR_DQ_FULL + X]);
            for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_Q_EMPTY]);      // the last readers of Q / dO (S, dP of tile nkv-1) are complete
          }
          __syncwarp();
          if (elect_one())
            for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_KV_EMPTY + st]);   // K(j): S(j), dQ(j);  V(j): dP(j)
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== softmax groups =====================
    reg_alloc<208>();
    const int X = (warp - 4) >> 2;            // 0: tile A, 1: tile B
    const int r = (tid - 128) & 127;          // row inside the tile = TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + X * kTileCols + lane_off, tdP = tS + 64, tdQ = tS + 128;
    const float c2 = a.scale * kLog2e;
    uint32_t steps = 0, done = 0;
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
      const int bh = item / a.pairs, qp = item - bh * a.pairs;
      const int h = bh % a.H, b = bh / a.H;
      const int q0 = qp * 2 * kTile + X * kTile;
      if (q0 >= T) continue;                  // tile B of the last pair may be empty (never tile A)
      const int row = q0 + r;
      // delta_r = sum_d dO[r, d] O[r, d];  log2-domain lse; both also go to the workspace of the dK / dV kernel (+inf / 0 in the padding)
      float delta = 0.f, lse2 = INFINITY;
      if (row < T) {
        const uint4* po = reinterpret_cast<const uint4*>(a.out + ((size_t)b * T + row) * a.ld_out + h * kD);
        const uint4* pd = reinterpret_cast<const uint4*>(a.dout + ((size_t)b * T + row) * a.ld_dout + h * kD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 x = po[c], y = pd[c];
          const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[u]));
            const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[u]));
            delta = fmaf(fx.x, fy.x, delta);
            delta = fmaf(fx.y, fy.y, delta);
          }
        }
        lse2 = a.lse[(size_t)bh * T + row] * kLog2e;
      }
      if (row < a.Tpad) {
        a.stats[(size_t)bh * a.Tpad + row] = delta;
        a.stats[(size_t)a.B * a.H * a.Tpad + (size_t)bh * a.Tpad + row] = lse2;
      }
      const float2 c2v = make_float2(c2, c2), nl = make_float2(-lse2, -lse2), nd = make_float2(-delta, -delta);
      for (int j = 0; j < nkv; ++j, ++steps) {
        mbar_wait(&bars[BAR_S_FULL + X], steps & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {           // 32 scores at a time: dS columns 16c .. 16c+15 go over S columns this thread has already read
          float s[32], dp[32];
          tmem_ld_32x32(tS + 32 * c, s);
          tmem_ld_32x32(tdP + 32 * c, dp);
          tc_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float2 e = ffma2(make_float2(s[i], s[i + 1]), c2v, nl);
            e.x = fast_ex2(e.x);
            e.y = fast_ex2(e.y);
            const float2 ds = fmul2(e, fadd2(make_float2(dp[i], dp[i + 1]), nd));   // dS = P (dP - delta)
            pk[i >> 1] = pack_bf16x2(ds.x, ds.y);
          }
          if (c == 1) {
            // chunk 1's dS lands in columns [16, 32): chunk 0 (columns [0, 32) of S) has been read, and so has chunk 1's own S ([32, 64))
          }
          tmem_st_32x16(tS + 16 * c, pk);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_DS_FULL + X]);
      }
      // ---- epilogue: dQ * scale -> bf16
      mbar_wait(&bars[BAR_DQ_FULL + X], done & 1);
      ++done;
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld_32x32(tdQ + 32 * c, v);
        tc_wait_ld();
        if (row < T) {
          uint4* dst = reinterpret_cast<uint4*>(a.dqkv + ((size_t)b * T + row) * a.ld_dqkv + h * kD + 32 * c);
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
      // the next item's dQ(0) (accumulate = 0) is gated by this group's next DS_FULL arrival, i.e. after these reads: no extra barrier
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
constexpr int kStages = 6;
constexpr int kStatBytes = 2 * kStep * 4;                         // delta[64], lse2[64] of the streamed query tile
constexpr int kStageBytes = 2 * kStepBytes + kStatBytes;          // Q, dO, statistics
// TMEM columns per tile X: S^T at X*256 (P^T overwrites its first 32), dP^T at +64 (dS^T over its first 32), dV at +128, dK at +192
constexpr int kTileCols = 256;
enum { BAR_KV_FULL = 0, BAR_KV_EMPTY = 1, BAR_Q_FULL = 2, BAR_Q_EMPTY = BAR_Q_FULL + kStages, BAR_S_FULL = BAR_Q_EMPTY + kStages /*[X]*/,
       BAR_PS_FULL = BAR_S_FULL + 2 /*[X]*/, BAR_ACC_FULL = BAR_PS_FULL + 2 /*[X]*/, BAR_COUNT = BAR_ACC_FULL + 2 };
constexpr int kSmem = 4 * kTileBytes /*K A,B, V A,B*/ + kStages * kStageBytes + BAR_COUNT * 8 + 64 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
mhsa_bwd_dkv_ws_kernel(const __grid_constant__ CUtensorMap tma_kv, const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_do, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;                              // [2][16 KB]
  uint8_t* sV = sK + 2 * kTileBytes;               // [2][16 KB]
  uint8_t* sQ = sV + 2 * kTileBytes;               // [kStages][8 KB]
  uint8_t* sdO = sQ + kStages * kStepBytes;        // [kStages][8 KB]
  float* sStat = reinterpret_cast<float*>(sdO + kStages * kStepBytes);   // [kStages][delta 64 | lse2 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + kStages * kStatBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, dim = a.dim;
  const int nq = (T + kStep - 1) / kStep;

  if (tid == 0) {
    tma_prefetch_desc(&tma_kv);
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_do);
    for (int i = 0; i < BAR_COUNT; ++i) {
      int count = 1;
      if (i >= BAR_PS_FULL && i < BAR_PS_FULL + 2) count = 4;
      if (i == BAR_KV_EMPTY || (i >= BAR_Q_EMPTY && i < BAR_Q_EMPTY + kStages)) count = 2;
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
        const int bh = item / a.pairs, kp = item - bh * a.pairs;
        const int h = bh % a.H, b = bh / a.H;
        const int k0 = kp * 2 * kTile;
        const bool activeB = k0 + kTile < T;
        mbar_wait(&bars[BAR_KV_EMPTY], (work & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[BAR_KV_FULL], activeB ? 4 * kTileBytes : 2 * kTileBytes);
        tma_load_3d(sK, &tma_kv, &bars[BAR_KV_FULL], dim + h * kD, k0, b);
        tma_load_3d(sV, &tma_kv, &bars[BAR_KV_FULL], 2 * dim + h * kD, k0, b);
        if (activeB) {
          tma_load_3d(sK + kTileBytes, &tma_kv, &bars[BAR_KV_FULL], dim + h * kD, k0 + kTile, b);
          tma_load_3d(sV + kTileBytes, &tma_kv, &bars[BAR_KV_FULL], 2 * dim + h * kD, k0 + kTile, b);
        }
        const float* g_delta = a.stats + (size_t)bh * a.Tpad;
        const float* g_lse2 = a.stats + (size_t)a.B * a.H * a.Tpad + (size_t)bh * a.Tpad;
        for (int i = 0; i < nq; ++i, ++q_iter) {
          const int st = q_iter % kStages;
          mbar_wait(&bars[BAR_Q_EMPTY + st], ((q_iter / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[BAR_Q_FULL + st], kStageBytes);
          tma_load_3d(sQ + st * kStepBytes, &tma_q, &bars[BAR_Q_FULL + st], h * kD, i * kStep, b);
          tma_load_3d(sdO + st * kStepBytes, &tma_do, &bars[BAR_Q_FULL + st], h * kD, i * kStep, b);
          bulk_load(sStat + st * 2 * kStep, g_delta + i * kStep, kStep * 4, &bars[BAR_Q_FULL + st]);
          bulk_load(sStat + st * 2 * kStep + kStep, g_lse2 + i * kStep, kStep * 4, &bars[BAR_Q_FULL + st]);
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers =====================
      const int X = warp - 1;
      const uint32_t k_addr = smem_u32(sK + X * kTileBytes), v_addr = smem_u32(sV + X * kTileBytes);
      const uint32_t tS = tmem + X * kTileCols, tdP = tS + 64, tdV = tS + 128, tdK = tS + 192;
      uint32_t q_base = 0, work = 0, steps = 0;
      auto issue_s = [&](uint32_t it) {   // S^T = K Q(it)^T and dP^T = V dO(it)^T
        const int st = it % kStages;
        mbar_wait(&bars[BAR_Q_FULL + st], (it / kStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          mma_kk(tS, k_addr, smem_u32(sQ + st * kStepBytes), false);
          mma_kk(tdP, v_addr, smem_u32(sdO + st * kStepBytes), false);
          umma_commit(&bars[BAR_S_FULL + X]);
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, q_base += nq) {
        const int kp = item % a.pairs;
        const bool activeB = kp * 2 * kTile + kTile < T;
        const int releases = (X == 0 && !activeB) ? 2 : 1;
        mbar_wait(&bars[BAR_KV_FULL], work & 1);
        tc_fence_after();
        if (X == 1 && !activeB) continue;
        issue_s(q_base);
        for (int i = 0; i < nq; ++i, ++steps) {
          const uint32_t iti = q_base + i;
          const int st = iti % kStages;
          mbar_wait(&bars[BAR_PS_FULL + X], steps & 1);
          tc_fence_after();
          if (elect_one()) {
            mma_tmn(tdV, tS, smem_u32(sdO + st * kStepBytes), i > 0);     // dV (+)= P^T dO(i)
            mma_tmn(tdK, tdP, smem_u32(sQ + st * kStepBytes), i > 0);     // dK (+)= dS^T Q(i)
          }
          __syncwarp();
          if (i + 1 < nq) {
            issue_s(iti + 1);
          } else if (elect_one()) {
            umma_commit(&bars[BAR_ACC_FULL + X]);
            for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_KV_EMPTY]);
          }
          __syncwarp();
          if (elect_one())
            for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_Q_EMPTY + st]);
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== softmax groups: thread = key row =====================
    reg_alloc<208>();
    const int X = (warp - 4) >> 2;
    const int r = (tid - 128) & 127;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + X * kTileCols + lane_off, tdP = tS + 64, tdV = tS + 128, tdK = tS + 192;
    const float c2 = a.scale * kLog2e;
    const float2 c2v = make_float2(c2, c2);
    uint32_t steps = 0, done = 0, q_base = 0;
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, q_base += nq) {
      const int bh = item / a.pairs, kp = item - bh * a.pairs;
      const int h = bh % a.H, b = bh / a.H;
      const int k0 = kp * 2 * kTile + X * kTile;
      if (k0 >= T) continue;
      for (int i = 0; i < nq; ++i, ++steps) {
        const int st = (q_base + i) % kStages;
        const float* s_delta = sStat + st * 2 * kStep;
        const float* s_lse2 = s_delta + kStep;
        mbar_wait(&bars[BAR_S_FULL + X], steps & 1);    // S_FULL completes after the stage's TMA / bulk copies: the statistics are visible
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float s[32], dp[32];
          tmem_ld_32x32(tS + 32 * c, s);
          tmem_ld_32x32(tdP + 32 * c, dp);
          tc_wait_ld();
          uint32_t pp[16], pd[16];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 l4 = *reinterpret_cast<const float4*>(s_lse2 + 32 * c + 4 * q4);
            const float4 d4 = *reinterpret_cast<const float4*>(s_delta + 32 * c + 4 * q4);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int q = 4 * q4 + 2 * u;
              const float2 ls = u == 0 ? make_float2(-l4.x, -l4.y) : make_float2(-l4.z, -l4.w);
              const float2 dl = u == 0 ? make_float2(-d4.x, -d4.y) : make_float2(-d4.z, -d4.w);
              float2 e = ffma2(make_float2(s[q], s[q + 1]), c2v, ls);
              e.x = fast_ex2(e.x);
              e.y = fast_ex2(e.y);
              const float2 ds = fmul2(e, fadd2(make_float2(dp[q], dp[q + 1]), dl));
              pp[q >> 1] = pack_bf16x2(e.x, e.y);       // P^T
              pd[q >> 1] = pack_bf16x2(ds.x, ds.y);     // dS^T
            }
          }
          tmem_st_32x16(tS + 16 * c, pp);
          tmem_st_32x16(tdP + 16 * c, pd);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_PS_FULL + X]);
      }
      // ---- epilogue: dK * scale, dV -> bf16
      mbar_wait(&bars[BAR_ACC_FULL + X], done & 1);
      ++done;
      tc_fence_after();
      const int row = k0 + r;
      __nv_bfloat16* base = a.dqkv + ((size_t)b * T + row) * a.ld_dqkv + h * kD;
#pragma unroll
      for (int w = 0; w < 2; ++w) {              // 0: dK, 1: dV
        const float sc = w == 0 ? a.scale : 1.0f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld_32x32((w == 0 ? tdK : tdV) + 32 * c, v);
          tc_wait_ld();
          if (row < T) {
            uint4* dst = reinterpret_cast<uint4*>(base + (w == 0 ? dim : 2 * dim) + 32 * c);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}
}  // namespace dkv
}  // namespace wsb

size_t mhsa_bwd_ws_floats(int B, int T, int H) {
  const size_t tpad = (size_t)(T + 127) / 128 * 128;   // the pipelined kernels pad to their 128-row blocks (covers the 64-row padding of this file's kernels)
  return 2 * (size_t)B * H * tpad;
}

int mhsa_bwd_ws(const gvk_mhsa_bwd_params* p, cudaStream_t stream) {
  using namespace wsb;
  static bool configured = false;
  if (!configured) {
    int st = cuda_status(cudaFuncSetAttribute(dq::mhsa_bwd_dq_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dq::kSmem), "mhsa_bwd_dq_ws smem");
    if (st != GVK_OK) return st;
    st = cuda_status(cudaFuncSetAttribute(dkv::mhsa_bwd_dkv_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv::kSmem), "mhsa_bwd_dkv_ws smem");
    if (st != GVK_OK) return st;
    configured = true;
  }
  const int dim = p->H * kD;
  CUtensorMap tq128, tq64, tdo128, tdo64;
  int st = make_tma_3d_bf16(&tq128, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, kTile, kD);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tq64, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, kStep, kD);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tdo128, p->dout, p->B, p->T, dim, p->ld_dout, (uint64_t)p->T * p->ld_dout, kTile, kD);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tdo64, p->dout, p->B, p->T, dim, p->ld_dout, (uint64_t)p->T * p->ld_dout, kStep, kD);
  if (st != GVK_OK) return st;
  Args a;
  a.B = p->B; a.T = p->T; a.H = p->H; a.dim = dim; a.scale = p->scale;
  a.Tpad = (p->T + kStep - 1) / kStep * kStep;
  a.out = reinterpret_cast<const __nv_bfloat16*>(p->out);
  a.ld_out = p->ld_out;
  a.dout = reinterpret_cast<const __nv_bfloat16*>(p->dout);
  a.ld_dout = p->ld_dout;
  a.lse = p->lse;
  a.stats = p->delta;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(p->dqkv);
  a.ld_dqkv = p->ld_dqkv;
  a.pairs = ((p->T + kTile - 1) / kTile + 1) / 2;
  a.num_items = p->B * p->H * a.pairs;
  const int grid = std::min(a.num_items, sm_count());
  dq::mhsa_bwd_dq_ws_kernel<<<grid, kThreads, dq::kSmem, stream>>>(tq128, tq64, tdo128, a);
  GVK_CHECK_LAUNCH("mhsa_bwd_dq_ws");
  dkv::mhsa_bwd_dkv_ws_kernel<<<grid, kThreads, dkv::kSmem, stream>>>(tq128, tq64, tdo64, a);
  GVK_CHECK_LAUNCH("mhsa_bwd_dkv_ws");
  return GVK_OK;
}

}  // namespace gvk
