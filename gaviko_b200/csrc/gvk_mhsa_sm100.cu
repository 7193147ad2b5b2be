// gvk_mhsa_sm100.cu — flash-style multi-head self-attention on the 5th-gen tensor cores (see include/gvk.h: gvk_mhsa_fwd / _bwd).
//
// One CTA = 128 threads = one 128-row tile of one (volume, head).  Thread t owns row t of the tile (TMEM lane t), so the row-wise
// softmax needs no cross-thread reductions.  Operand tiles are TMA loads with SWIZZLE_128B; S = Q K^T, P V and the backward
// products are tcgen05.mma with fp32 accumulators in TMEM; P / dS go registers -> swizzled smem -> A operand of the next MMA.
// "Transposed" products (P V, dS K, P^T dO, dS^T Q) read their [rows x 64] TMA tiles as MN-major B operands, so no tensor is
// ever transposed in memory.  Two or three CTAs are resident per SM, which is what overlaps one CTA's softmax with another's MMAs.
//
//   forward  : grid (ceil(T/128), B*H); loop over 64-row K/V tiles; online softmax; O accumulated in registers.
//   backward : dQ kernel  — grid (ceil(T/128), B*H) over query tiles, loops over 64-row K/V tiles (also writes delta = rowsum(dO*O));
//              dKV kernel — grid (ceil(T/128), B*H) over key tiles, loops over 64-row Q/dO tiles.
//              Two kernels instead of one fused kernel: no atomics on dQ, deterministic results.
// Replaces reference model/vision_transformer.py:65-71 (+ its autograd backward), which materialises (B, H, T, T) twice.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "gvk_common.cuh"

namespace gvk {

constexpr int kFaThreads = 128;
constexpr int kFaD = 64;               // head dim
constexpr int kTileBytes128 = 128 * 128;  // [128 rows x 64 bf16]
constexpr int kTileBytes64 = 64 * 128;    // [ 64 rows x 64 bf16]
constexpr float kLog2e = 1.4426950408889634f;

struct FaDesc {  // MN-major descriptor fields (debug-tunable through GVK_FA_DESC="lbo,sbo,kadv")
  uint32_t lbo, sbo, kadv;
  uint32_t tmem_p;   // 1: P / dS go registers -> TMEM and feed the next MMA as its A operand (no shared-memory round trip)
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Write 64 consecutive values of row r into a [rows x 64] bf16 K-major SWIZZLE_128B tile (the layout TMA would produce).
__device__ __forceinline__ void store_row_sw128(uint8_t* tile, int r, const float (&v)[64]) {
  uint8_t* row = tile + r * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 pk;
    pk.x = pack_bf16(v[8 * c + 0], v[8 * c + 1]);
    pk.y = pack_bf16(v[8 * c + 2], v[8 * c + 3]);
    pk.z = pack_bf16(v[8 * c + 4], v[8 * c + 5]);
    pk.w = pack_bf16(v[8 * c + 6], v[8 * c + 7]);
    *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = pk;
  }
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  tmem_ld_32x32(taddr, *reinterpret_cast<float(*)[32]>(&v[0]));
  tmem_ld_32x32(taddr + 32, *reinterpret_cast<float(*)[32]>(&v[32]));
}
// D[tmem, 128 x 64] (+)= A[128 x 64 K-major tile] * B[64 x 64 K-major tile]^T
__device__ __forceinline__ void mma_kk(uint32_t d_tmem, const void* a, const void* b, bool accumulate) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  const uint32_t a_addr = smem_u32(a), b_addr = smem_u32(b);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16(d_tmem, make_sw128_desc(a_addr + k * 32, 16, 1024), make_sw128_desc(b_addr + k * 32, 16, 1024), idesc, (accumulate || k > 0) ? 1u : 0u);
}
// D[tmem, 128 x 64] (+)= A[128 x 64 K-major tile] * B, with B given as a [64 (k) rows x 64 (n)] tile (n contiguous): MN-major B
__device__ __forceinline__ void mma_kmn(uint32_t d_tmem, const void* a, const void* b, bool accumulate, const FaDesc& fd) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
  const uint32_t a_addr = smem_u32(a), b_addr = smem_u32(b);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16(d_tmem, make_sw128_desc(a_addr + k * 32, 16, 1024), make_sw128_desc(b_addr + k * fd.kadv, fd.lbo, fd.sbo), idesc, (accumulate || k > 0) ? 1u : 0u);
}

// D[tmem, 128 x 64] (+)= A * B with A = [128 x 64] bf16 in TMEM (32 columns starting at a_tmem) and B a [64 (k) x 64 (n)] MN-major smem tile
__device__ __forceinline__ void mma_tmn(uint32_t d_tmem, uint32_t a_tmem, const void* b, bool accumulate, const FaDesc& fd) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
  const uint32_t b_addr = smem_u32(b);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16_ts(d_tmem, a_tmem + 8 * k, make_sw128_desc(b_addr + k * fd.kadv, fd.lbo, fd.sbo), idesc, (accumulate || k > 0) ? 1u : 0u);
}
// Row of 64 probabilities -> bf16 pairs -> 32 TMEM columns of this thread's lane
__device__ __forceinline__ void store_row_tmem(uint32_t taddr, const float (&v)[64]) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
  tmem_st_32x32(taddr, r);
  tc_wait_st();
}

struct FaCommon {
  int B, T, H, dim;
  float scale;
};

// =================================================================================================
// forward
// =================================================================================================
struct FwdArgs {
  FaCommon c;
  __nv_bfloat16* out;
  int ld_out;
  float* lse;
  FaDesc fd;
};
constexpr int kKvStages = 3;   // K/V (or Q/dO) ring depth: a TMA tile must be requested >= 2 iterations (~1.5 us) ahead to hide the L2 -> smem latency
constexpr int kFwdSmem = kTileBytes128 /*Q*/ + 2 * kKvStages * kTileBytes64 /*K,V ring*/ + kTileBytes128 /*P*/ + 128 + 1024;

__global__ void __launch_bounds__(kFaThreads, 3)
mhsa_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv, FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB alignment by pointer arithmetic: keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes128;                    // kKvStages stages
  uint8_t* sV = sK + kKvStages * kTileBytes64;         // kKvStages stages
  uint8_t* sP = sV + kKvStages * kTileBytes64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kTileBytes128);  // q, s, pv, kv[kKvStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128;
  const int bh = blockIdx.y, h = bh % a.c.H, b = bh / a.c.H;
  const int T = a.c.T, dim = a.c.dim;

  if (tid == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    for (int i = 0; i < 3 + kKvStages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_pv = tmem + 64;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  uint64_t *bar_q = &bars[0], *bar_s = &bars[1], *bar_pv = &bars[2], *bar_kv = &bars[3];

  const int nkv = (T + 63) / 64;
  auto load_kv = [&](int j) {   // tid 0 only
    const int st = j % kKvStages;
    mbar_arrive_expect_tx(&bar_kv[st], 2 * kTileBytes64);
    tma_load_3d(sK + st * kTileBytes64, &tma_kv, &bar_kv[st], dim + h * kFaD, j * 64, b);
    tma_load_3d(sV + st * kTileBytes64, &tma_kv, &bar_kv[st], 2 * dim + h * kFaD, j * 64, b);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, kTileBytes128);
    tma_load_3d(sQ, &tma_q, bar_q, h * kFaD, q0, b);
    for (int j = 0; j < kKvStages - 1 && j < nkv; ++j) load_kv(j);
  }
  const float c2 = a.c.scale * kLog2e;
  float m = -INFINITY, l = 0.f;
  float o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
  mbar_wait(bar_q, 0);

  for (int j = 0; j < nkv; ++j) {
    const int buf = j % kKvStages;
    // stage (j + kKvStages - 1) % kKvStages was read by iteration j - 1, whose MMAs completed before its bar_pv wait returned
    if (tid == 0 && j + kKvStages - 1 < nkv) load_kv(j + kKvStages - 1);
    mbar_wait(&bar_kv[buf], (j / kKvStages) & 1);
    if (tid == 0) {
      tc_fence_after();
      mma_kk(t_s, sQ, sK + buf * kTileBytes64, false);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, j & 1);
    tc_fence_after();
    float s[64];
    tmem_ld64(t_s + lane_off, s);
    tc_wait_ld();
    const int valid = T - j * 64;  // columns >= valid are padding (zero-filled K rows)
    if (valid < 64) {
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i >= valid) s[i] = -INFINITY;
    }
    float mx = m;
#pragma unroll
    for (int i = 0; i < 64; ++i) mx = fmaxf(mx, s[i]);
    const float alpha = fast_ex2((m - mx) * c2);
    m = mx;
    const float mc = mx * c2;
    float rs = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      s[i] = fast_ex2(fmaf(s[i], c2, -mc));
      rs += s[i];
    }
    l = fmaf(l, alpha, rs);
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] *= alpha;
    if (a.fd.tmem_p) {
      store_row_tmem(t_s + lane_off, s);   // P overwrites the first 32 columns of S (this thread has already read its S row)
    } else {
      store_row_sw128(sP, tid, s);
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      if (a.fd.tmem_p) mma_tmn(t_pv, t_s, sV + buf * kTileBytes64, false, a.fd);
      else mma_kmn(t_pv, sP, sV + buf * kTileBytes64, false, a.fd);
      umma_commit(bar_pv);
    }
    mbar_wait(bar_pv, j & 1);
    tc_fence_after();
    tmem_ld64(t_pv + lane_off, s);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] += s[i];
  }
  const int row = q0 + tid;
  if (row < T) {
    const float inv = 1.0f / l;
    uint4* dst = reinterpret_cast<uint4*>(a.out + ((size_t)b * T + row) * a.ld_out + h * kFaD);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 pk;
      pk.x = pack_bf16(o[8 * c + 0] * inv, o[8 * c + 1] * inv);
      pk.y = pack_bf16(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
      pk.z = pack_bf16(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
      pk.w = pack_bf16(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
      dst[c] = pk;
    }
    a.lse[(size_t)bh * T + row] = m * a.c.scale + __logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// =================================================================================================
// backward: dQ (and delta)
// 256 threads per 128-row tile: thread (half, r) owns row r and score columns 32*half .. 32*half+31.  Four warps per tile left each
// sub-partition with two warps (two CTAs per SM) and the per-thread chain of 64 exp / multiply steps latency-bound at ~0.13 IPC per warp;
// eight warps halve the chain and double the warps that can hide TMEM / SFU latency.
// =================================================================================================
constexpr int kBwdThreads = 256;
struct BwdArgs {
  FaCommon c;
  const __nv_bfloat16* out;
  int ld_out;
  const __nv_bfloat16* dout;
  int ld_dout;
  const float* lse;
  float* delta;
  __nv_bfloat16* dqkv;
  int ld_dqkv;
  FaDesc fd;
};
// D[tmem, 128 x 64] (+)= A * B, A = [128 x 64] bf16 in TMEM as two 16-column pieces (k 0..31 at a_lo, k 32..63 at a_hi), B MN-major smem tile
__device__ __forceinline__ void mma_tmn_split(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, const void* b, bool accumulate, const FaDesc& fd) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
  const uint32_t b_addr = smem_u32(b);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16_ts(d_tmem, (k < 2 ? a_lo + 8 * k : a_hi + 8 * (k - 2)), make_sw128_desc(b_addr + k * fd.kadv, fd.lbo, fd.sbo), idesc, (accumulate || k > 0) ? 1u : 0u);
}
__device__ __forceinline__ void tmem_st_32x16b(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 32 values of this thread's row -> 16 packed bf16 pairs -> 16 TMEM columns
__device__ __forceinline__ void store_half_row_tmem(uint32_t taddr, const float (&v)[32]) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
  tmem_st_32x16b(taddr, r);
  tc_wait_st();
}
// 32 fp32 values -> bf16 -> 64 bytes of a global row
__device__ __forceinline__ void store_half_row_global(__nv_bfloat16* dst, const float (&v)[32], float scale) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 pk;
    pk.x = pack_bf16(v[8 * c + 0] * scale, v[8 * c + 1] * scale);
    pk.y = pack_bf16(v[8 * c + 2] * scale, v[8 * c + 3] * scale);
    pk.z = pack_bf16(v[8 * c + 4] * scale, v[8 * c + 5] * scale);
    pk.w = pack_bf16(v[8 * c + 6] * scale, v[8 * c + 7] * scale);
    d4[c] = pk;
  }
}
constexpr int kDqSmem = 2 * kTileBytes128 /*Q,dO*/ + 2 * kKvStages * kTileBytes64 /*K,V ring*/ + 128 + 1024;

__global__ void __launch_bounds__(kBwdThreads, 2)
mhsa_bwd_dq_sm100_kernel(const __grid_constant__ CUtensorMap tma_q128, const __grid_constant__ CUtensorMap tma_kv64, const __grid_constant__ CUtensorMap tma_do128,
                         BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB alignment by pointer arithmetic: keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + kTileBytes128;
  uint8_t* sK = sdO + kTileBytes128;
  uint8_t* sV = sK + kKvStages * kTileBytes64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKvStages * kTileBytes64);  // q, s, dq, kv[kKvStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int half = tid >> 7, r = tid & 127;
  const int q0 = blockIdx.x * 128;
  const int bh = blockIdx.y, h = bh % a.c.H, b = bh / a.c.H;
  const int T = a.c.T, dim = a.c.dim;

  if (tid == 0) {
    tma_prefetch_desc(&tma_q128);
    tma_prefetch_desc(&tma_kv64);
    tma_prefetch_desc(&tma_do128);
    for (int i = 0; i < 3 + kKvStages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_dp = tmem + 64, t_dq = tmem + 128;
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  uint64_t *bar_q = &bars[0], *bar_s = &bars[1], *bar_dq = &bars[2], *bar_kv = &bars[3];
  const int nkv = (T + 63) / 64;
  auto load_kv = [&](int j) {   // tid 0 only
    const int st = j % kKvStages;
    mbar_arrive_expect_tx(&bar_kv[st], 2 * kTileBytes64);
    tma_load_3d(sK + st * kTileBytes64, &tma_kv64, &bar_kv[st], dim + h * kFaD, j * 64, b);
    tma_load_3d(sV + st * kTileBytes64, &tma_kv64, &bar_kv[st], 2 * dim + h * kFaD, j * 64, b);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * kTileBytes128);
    tma_load_3d(sQ, &tma_q128, bar_q, h * kFaD, q0, b);
    tma_load_3d(sdO, &tma_do128, bar_q, h * kFaD, q0, b);
    for (int j = 0; j < kKvStages - 1 && j < nkv; ++j) load_kv(j);
  }
  // delta_r = sum_d dO[r,d] * O[r,d] (both halves compute it; half 0 publishes it together with the log2-domain lse)
  const int row = q0 + r;
  float delta = 0.f, lse2 = 0.f;
  if (row < T) {
    const uint4* po = reinterpret_cast<const uint4*>(a.out + ((size_t)b * T + row) * a.ld_out + h * kFaD);
    const uint4* pd = reinterpret_cast<const uint4*>(a.dout + ((size_t)b * T + row) * a.ld_dout + h * kFaD);
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
    if (half == 0) {
      a.delta[(size_t)bh * T + row] = delta;
      a.delta[(size_t)a.c.B * a.c.H * T + (size_t)bh * T + row] = lse2;   // second half of the workspace: log2-domain lse for the dK/dV kernel
    }
  }
  const float c2 = a.c.scale * kLog2e;
  // this thread's dS piece goes over S columns it has itself consumed: [0,16) for half 0, [32,48) for half 1
  const uint32_t t_s_mine = t_s + 32 * half + lane_off, t_dp_mine = t_dp + 32 * half + lane_off;
  mbar_wait(bar_q, 0);
  for (int j = 0; j < nkv; ++j) {
    const int buf = j % kKvStages;
    if (tid == 0 && j + kKvStages - 1 < nkv) {
      // the stage being refilled was last read by the MMAs of iteration j - 1: wait for them here (thread 0 only) instead of stalling every
      // thread at the end of each iteration — tcgen05.mma executes in issue order, so S(j) cannot overtake dQ(j-1)'s read of dS(j-1)
      if (j > 0) mbar_wait(bar_dq, (j - 1) & 1);
      load_kv(j + kKvStages - 1);
    }
    mbar_wait(&bar_kv[buf], (j / kKvStages) & 1);
    if (warp == 0) {   // warp-uniform control flow around the issue: only the tcgen05 instructions sit under elect_one
      tc_fence_after();
      if (elect_one()) {
        mma_kk(t_s, sQ, sK + buf * kTileBytes64, false);      // S  = Q K^T
        mma_kk(t_dp, sdO, sV + buf * kTileBytes64, false);    // dP = dO V^T
        umma_commit(bar_s);
      }
      __syncwarp();
    }
    mbar_wait(bar_s, j & 1);
    tc_fence_after();
    float s[32], dp[32];
    tmem_ld_32x32(t_s_mine, s);
    tmem_ld_32x32(t_dp_mine, dp);
    tc_wait_ld();
    const int valid = T - j * 64 - 32 * half;   // columns >= valid of this thread's 32 are zero-filled padding
    if (valid >= 32) {   // packed fp32 math (FFMA2 / FADD2 / FMUL2): 3 issue slots per score instead of 4.5
      const float2 c2v = make_float2(c2, c2), nl = make_float2(-lse2, -lse2), nd = make_float2(-delta, -delta);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float2 e = ffma2(make_float2(s[i], s[i + 1]), c2v, nl);
        e.x = fast_ex2(e.x);
        e.y = fast_ex2(e.y);
        const float2 ds = fmul2(e, fadd2(make_float2(dp[i], dp[i + 1]), nd));   // dS = P (dP - delta)
        s[i] = ds.x;
        s[i + 1] = ds.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p = (i < valid) ? fast_ex2(fmaf(s[i], c2, -lse2)) : 0.f;
        s[i] = p * (dp[i] - delta);
      }
    }
    store_half_row_tmem(t_s_mine, s);   // dS (bf16 pairs): the TMEM-resident A operand of the next MMA
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_tmn_split(t_dq, t_s, t_s + 32, sK + buf * kTileBytes64, j > 0, a.fd);   // dQ += dS K
        umma_commit(bar_dq);
      }
      __syncwarp();
    }
  }
  mbar_wait(bar_dq, (nkv - 1) & 1);
  tc_fence_after();
  {
    float dq[32];
    tmem_ld_32x32(t_dq + 32 * half + lane_off, dq);
    tc_wait_ld();
    if (row < T) store_half_row_global(a.dqkv + ((size_t)b * T + row) * a.ld_dqkv + h * kFaD + 32 * half, dq, a.c.scale);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// =================================================================================================
// backward: dK, dV  (same thread layout: half = 32-column block of the 64 queries of the tile in flight)
// =================================================================================================
constexpr int kDkvStages = 3;
constexpr int kDkvSmem = 2 * kTileBytes128 /*K,V*/ + 2 * kDkvStages * kTileBytes64 /*Q,dO ring*/ + 2 * kDkvStages * 64 * 4 /*lse, delta ring*/ + 128 + 1024;

__global__ void __launch_bounds__(kBwdThreads, 2)
mhsa_bwd_dkv_sm100_kernel(const __grid_constant__ CUtensorMap tma_kv128, const __grid_constant__ CUtensorMap tma_q64, const __grid_constant__ CUtensorMap tma_do64,
                          BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB alignment by pointer arithmetic: keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTileBytes128;
  uint8_t* sQ = sV + kTileBytes128;                     // kDkvStages stages of [64 x 64]
  uint8_t* sdO = sQ + kDkvStages * kTileBytes64;        // kDkvStages stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdO + kDkvStages * kTileBytes64);    // kv, s, acc, q[kDkvStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_lse = reinterpret_cast<float*>(bars + 16);    // [kDkvStages][64] log2-domain lse of the q tile in flight (16-byte aligned)
  float* s_delta = s_lse + kDkvStages * 64;              // [kDkvStages][64]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int half = tid >> 7, r = tid & 127;
  const int k0 = blockIdx.x * 128;
  const int bh = blockIdx.y, h = bh % a.c.H, b = bh / a.c.H;
  const int T = a.c.T, dim = a.c.dim;

  if (tid == 0) {
    tma_prefetch_desc(&tma_kv128);
    tma_prefetch_desc(&tma_q64);
    tma_prefetch_desc(&tma_do64);
    for (int i = 0; i < 3 + kDkvStages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  const int nq = (T + 63) / 64;
  // lse / delta of q tile i travel with cp.async two iterations ahead (a blocking global load here sat on every iteration's critical
  // path); rows past T read a clamped (finite) element and are masked by `valid` below
  auto load_stats = [&](int i) {   // threads 0..63; always commits a group so the wait_group accounting stays uniform
    if (i < nq) {
      const size_t q = (size_t)bh * T + min(i * 64 + tid, T - 1);
      cp_async4(&s_lse[(i % kDkvStages) * 64 + tid], a.delta + (size_t)a.c.B * a.c.H * T + q);   // log2-domain lse written by the dQ kernel
      cp_async4(&s_delta[(i % kDkvStages) * 64 + tid], a.delta + q);
    }
    cp_async_commit();
  };
  if (tid < 64) {
    for (int i = 0; i < kDkvStages - 1; ++i) load_stats(i);
    cp_async_wait<kDkvStages - 2>();   // tile 0 landed; made visible to everyone by the __syncthreads below
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_dp = tmem + 64, t_dv = tmem + 128, t_dk = tmem + 192;
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  uint64_t *bar_kv = &bars[0], *bar_s = &bars[1], *bar_acc = &bars[2], *bar_q = &bars[3];
  auto load_q = [&](int i) {   // tid 0 only
    const int st = i % kDkvStages;
    mbar_arrive_expect_tx(&bar_q[st], 2 * kTileBytes64);
    tma_load_3d(sQ + st * kTileBytes64, &tma_q64, &bar_q[st], h * kFaD, i * 64, b);
    tma_load_3d(sdO + st * kTileBytes64, &tma_do64, &bar_q[st], h * kFaD, i * 64, b);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_kv, 2 * kTileBytes128);
    tma_load_3d(sK, &tma_kv128, bar_kv, dim + h * kFaD, k0, b);
    tma_load_3d(sV, &tma_kv128, bar_kv, 2 * dim + h * kFaD, k0, b);
    for (int i = 0; i < kDkvStages - 1 && i < nq; ++i) load_q(i);
  }
  const float c2 = a.c.scale * kLog2e;
  const bool row_valid = (k0 + r) < T;
  const uint32_t t_s_mine = t_s + 32 * half + lane_off, t_dp_mine = t_dp + 32 * half + lane_off;
  mbar_wait(bar_kv, 0);
  for (int i = 0; i < nq; ++i) {
    const int buf = i % kDkvStages;
    if (tid == 0 && i + kDkvStages - 1 < nq) {
      if (i > 0) mbar_wait(bar_acc, (i - 1) & 1);     // the stage being refilled was last read by the MMAs of iteration i - 1
      load_q(i + kDkvStages - 1);
    }
    if (tid < 64) load_stats(i + kDkvStages - 1);
    mbar_wait(&bar_q[buf], (i / kDkvStages) & 1);
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_kk(t_s, sK, sQ + buf * kTileBytes64, false);      // S^T  = K Q^T
        mma_kk(t_dp, sV, sdO + buf * kTileBytes64, false);    // dP^T = V dO^T
        umma_commit(bar_s);
      }
      __syncwarp();
    }
    mbar_wait(bar_s, i & 1);
    tc_fence_after();
    float s[32], dp[32];
    tmem_ld_32x32(t_s_mine, s);
    tmem_ld_32x32(t_dp_mine, dp);
    tc_wait_ld();
    const int valid = row_valid ? T - i * 64 - 32 * half : 0;
    const float4* lse4 = reinterpret_cast<const float4*>(s_lse + buf * 64 + 32 * half);
    const float4* del4 = reinterpret_cast<const float4*>(s_delta + buf * 64 + 32 * half);
    if (valid >= 32) {                     // packed fp32 math, two query columns per instruction
      const float2 c2v = make_float2(c2, c2);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const float4 l4 = lse4[q4], d4 = del4[q4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int q = 4 * q4 + 2 * u;
          const float2 ls = u == 0 ? make_float2(l4.x, l4.y) : make_float2(l4.z, l4.w);
          const float2 dl = u == 0 ? make_float2(d4.x, d4.y) : make_float2(d4.z, d4.w);
          float2 e = fadd2(fmul2(make_float2(s[q], s[q + 1]), c2v), make_float2(-ls.x, -ls.y));
          e.x = fast_ex2(e.x);
          e.y = fast_ex2(e.y);
          const float2 ds = fmul2(e, fadd2(make_float2(dp[q], dp[q + 1]), make_float2(-dl.x, -dl.y)));
          s[q] = e.x; s[q + 1] = e.y;            // P^T
          dp[q] = ds.x; dp[q + 1] = ds.y;        // dS^T
        }
      }
    } else {                               // last query tile / key rows past T
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const float4 l4 = lse4[q4], d4 = del4[q4];
        const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = 4 * q4 + u;
          const float p = (q < valid) ? fast_ex2(fmaf(s[q], c2, -ls[u])) : 0.f;
          s[q] = p;
          dp[q] = p * (dp[q] - dl[u]);
        }
      }
    }
    store_half_row_tmem(t_s_mine, s);     // P^T  over the S^T  columns this thread consumed
    store_half_row_tmem(t_dp_mine, dp);   // dS^T over the dP^T columns this thread consumed
    if (tid < 64) cp_async_wait<kDkvStages - 2>();   // stats of tile i + 1 landed: visible after this barrier
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_tmn_split(t_dv, t_s, t_s + 32, sdO + buf * kTileBytes64, i > 0, a.fd);    // dV += P^T dO
        mma_tmn_split(t_dk, t_dp, t_dp + 32, sQ + buf * kTileBytes64, i > 0, a.fd);   // dK += dS^T Q
        umma_commit(bar_acc);
      }
      __syncwarp();
    }
  }
  mbar_wait(bar_acc, (nq - 1) & 1);
  tc_fence_after();
  {
    float v[32];
    const int row = k0 + r;
    __nv_bfloat16* base = a.dqkv + ((size_t)b * T + row) * a.ld_dqkv + h * kFaD + 32 * half;
    tmem_ld_32x32(t_dk + 32 * half + lane_off, v);
    tc_wait_ld();
    if (row < T) store_half_row_global(base + dim, v, a.c.scale);
    tmem_ld_32x32(t_dv + 32 * half + lane_off, v);
    tc_wait_ld();
    if (row < T) store_half_row_global(base + 2 * dim, v, 1.0f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// =================================================================================================
// host
// =================================================================================================
static FaDesc fa_desc() {
  static FaDesc fd = {8192, 1024, 2048, 0};
  static bool init = false;
  if (!init) {
    init = true;
    if (const char* e = getenv("GVK_FA_DESC")) {
      unsigned a, b, c;
      if (sscanf(e, "%u,%u,%u", &a, &b, &c) == 3) fd = {a, b, c, fd.tmem_p};
    }
    if (const char* e = getenv("GVK_FA_TMEM_P")) fd.tmem_p = atoi(e) != 0;
  }
  return fd;
}

template <typename K>
static int set_smem(K kern, int bytes, const char* what) {
  return cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), what);
}

static int check_common(const void* qkv, int ld, int B, int T, int H, const char* who) {
  GVK_CHECK_ARG(qkv && B > 0 && T > 0 && H > 0, "%s: bad argument", who);
  GVK_CHECK_ARG(ld % 8 == 0 && ld >= 3 * H * kFaD, "%s: ld=%d must be a multiple of 8 and >= 3*H*64", who, ld);
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "%s: qkv must be 16-byte aligned", who);
  GVK_CHECK_ARG((long long)B * H <= 65535, "%s: B*H=%lld exceeds the grid limit", who, (long long)B * H);
  return GVK_OK;
}

int mhsa_fwd(const gvk_mhsa_fwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->out && p->lse, "gvk_mhsa_fwd: null pointer");
  int st = check_common(p->qkv, p->ld, p->B, p->T, p->H, "gvk_mhsa_fwd");
  if (st != GVK_OK) return st;
  GVK_CHECK_ARG(p->ld_out % 8 == 0, "gvk_mhsa_fwd: ld_out must be a multiple of 8");
  GVK_CHECK_ARG(p->drop_p >= 0.f && p->drop_p < 1.f, "gvk_mhsa_fwd: drop_p must be in [0, 1)");
  {
    // GVK_MHSA_IMPL: 2 = 64-key-step warp-specialised kernel (default: 420 us at B = 64), 3 = 128-key-step kernel with the SFU token (440 us; see its
    // header for what the timeline showed), 1 = one tile per CTA (A/B comparison)
    static int impl = -1;
    if (impl < 0) { const char* e = getenv("GVK_MHSA_IMPL"); impl = e ? atoi(e) : 2; }
    if (impl == 3) return mhsa_fwd2(p, stream);
    if (impl == 2) return mhsa_ws_fwd(p, stream);
  }
  GVK_CHECK_ARG(p->drop_p == 0.f, "gvk_mhsa_fwd: attention dropout needs the warp-specialised kernel (GVK_MHSA_IMPL unset)");
  static bool configured = false;
  if (!configured) {
    st = set_smem(mhsa_fwd_sm100_kernel, kFwdSmem, "mhsa_fwd smem");
    if (st != GVK_OK) return st;
    configured = true;
  }
  const int dim = p->H * kFaD;
  CUtensorMap tq, tkv;
  st = make_tma_3d_bf16(&tq, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, 128, 64);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tkv, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, 64, 64);
  if (st != GVK_OK) return st;
  FwdArgs a;
  a.c = {p->B, p->T, p->H, dim, p->scale};
  a.out = reinterpret_cast<__nv_bfloat16*>(p->out);
  a.ld_out = p->ld_out;
  a.lse = p->lse;
  a.fd = fa_desc();
  dim3 grid((p->T + 127) / 128, p->B * p->H);
  mhsa_fwd_sm100_kernel<<<grid, kFaThreads, kFwdSmem, stream>>>(tq, tkv, a);
  GVK_CHECK_LAUNCH("mhsa_fwd_sm100");
  return GVK_OK;
}

int mhsa_bwd(const gvk_mhsa_bwd_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->out && p->lse && p->dout && p->delta && p->dqkv, "gvk_mhsa_bwd: null pointer");
  int st = check_common(p->qkv, p->ld, p->B, p->T, p->H, "gvk_mhsa_bwd");
  if (st != GVK_OK) return st;
  GVK_CHECK_ARG(p->ld_out % 8 == 0 && p->ld_dout % 8 == 0 && p->ld_dqkv % 8 == 0, "gvk_mhsa_bwd: leading dimensions must be multiples of 8");
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(p->dout) & 15) == 0, "gvk_mhsa_bwd: dout must be 16-byte aligned");
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(p->delta) & 15) == 0, "gvk_mhsa_bwd: the workspace must be 16-byte aligned");
  {
    static int impl = -1;   // GVK_MHSA_BWD_IMPL: 3 = software-pipelined kernels (default), 2 = ping-pong kernels, 1 = one tile per CTA (A/B comparison)
    if (impl < 0) { const char* e = getenv("GVK_MHSA_BWD_IMPL"); impl = e ? atoi(e) : 3; }
    if (impl == 3) return mhsa_bwd_pipe(p, stream);
    GVK_CHECK_ARG(p->drop_p == 0.f || impl == 3, "gvk_mhsa_bwd: attention dropout needs the pipelined kernels (GVK_MHSA_BWD_IMPL unset)");
    if (impl == 2) return mhsa_bwd_ws(p, stream);
  }
  static bool configured = false;
  if (!configured) {
    st = set_smem(mhsa_bwd_dq_sm100_kernel, kDqSmem, "mhsa_bwd_dq smem");
    if (st != GVK_OK) return st;
    st = set_smem(mhsa_bwd_dkv_sm100_kernel, kDkvSmem, "mhsa_bwd_dkv smem");
    if (st != GVK_OK) return st;
    configured = true;
  }
  const int dim = p->H * kFaD;
  CUtensorMap tq128, tq64, tdo128, tdo64;
  st = make_tma_3d_bf16(&tq128, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, 128, 64);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tq64, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, 64, 64);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tdo128, p->dout, p->B, p->T, dim, p->ld_dout, (uint64_t)p->T * p->ld_dout, 128, 64);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tdo64, p->dout, p->B, p->T, dim, p->ld_dout, (uint64_t)p->T * p->ld_dout, 64, 64);
  if (st != GVK_OK) return st;
  BwdArgs a;
  a.c = {p->B, p->T, p->H, dim, p->scale};
  a.out = reinterpret_cast<const __nv_bfloat16*>(p->out);
  a.ld_out = p->ld_out;
  a.dout = reinterpret_cast<const __nv_bfloat16*>(p->dout);
  a.ld_dout = p->ld_dout;
  a.lse = p->lse;
  a.delta = p->delta;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(p->dqkv);
  a.ld_dqkv = p->ld_dqkv;
  a.fd = fa_desc();
  dim3 grid((p->T + 127) / 128, p->B * p->H);
  mhsa_bwd_dq_sm100_kernel<<<grid, kBwdThreads, kDqSmem, stream>>>(tq128, tq64, tdo128, a);
  GVK_CHECK_LAUNCH("mhsa_bwd_dq_sm100");
  mhsa_bwd_dkv_sm100_kernel<<<grid, kBwdThreads, kDkvSmem, stream>>>(tq128, tq64, tdo64, a);
  GVK_CHECK_LAUNCH("mhsa_bwd_dkv_sm100");
  return GVK_OK;
}

}  // namespace gvk
