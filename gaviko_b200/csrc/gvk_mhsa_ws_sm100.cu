// gvk_mhsa_ws_sm100.cu — warp-specialised flash attention for the frozen MHSA core (head dim 64, bf16 operands, fp32 softmax).
//
// One persistent CTA per SM, 12 warps:
//   warp 0      TMA producer   (Q tiles of the work item, then a ring of K/V tiles)
//   warp 1      MMA issuer     (one thread issues every tcgen05.mma; completion is signalled with tcgen05.commit -> mbarrier)
//   warps 4-7   softmax group A   (thread t owns query row t of tile A = TMEM lane t)
//   warps 8-11  softmax group B   (same for tile B; the two tiles ping-pong so the tensor pipe works on one tile's S / PV while the
//                                  other tile is in its exp2 phase — at head dim 64 the SFU, not the tensor pipe, is the bound)
// S = Q K^T lands in TMEM, is read once into registers (thread = row: no shuffles), P goes back to TMEM as packed bf16 over the columns
// S occupied and feeds O += P V as the TMEM-resident A operand (no shared-memory round trip); O stays in TMEM across the whole KV loop
// and is rescaled only when the running row maximum grows by more than 2^8 (exact: the final 1/l normalisation absorbs the stale max).
//
// Replaces model/vision_transformer.py:65-71 (softmax(q k^T * scale) v on the prompt-extended sequence); any T (tails masked in-kernel).
#include <algorithm>
#include <cstdlib>

#include "gvk_common.cuh"

namespace gvk {

namespace ws {
constexpr int kThreads = 384;
constexpr int kD = 64;
constexpr int kTile = 128;                   // query rows per tile
constexpr int kKV = 64;                      // keys per K/V tile
constexpr int kQBytes = kTile * kD * 2;      // 16 KB
constexpr int kKVBytes = kKV * kD * 2;       // 8 KB
constexpr int kStages = 8;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;    // log2 units
constexpr int kSBuf = 3;                     // score buffers per tile in TMEM: 2 tiles x (3 x 64 + 64 (O)) = 512 columns
constexpr int kOCol = 2 * kSBuf * 64;        // first O column

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

struct FwdArgs {
  int B, T, H, dim;
  float scale;
  __nv_bfloat16* out;
  int ld_out;
  float* lse;
  int pairs;       // query-tile pairs per (volume, head)
  int num_items;   // B * H * pairs
  uint32_t mn_lbo, mn_sbo, mn_kadv;   // MN-major B descriptor fields of the V tile
  int dbg;   // timing experiments only (GVK_WS_DBG): 1 = no exp2, 2 = no TMEM load of S; results are wrong
  MhsaDrop drop;   // attention-probability dropout (kDrop instantiation only)
};

// TMEM columns: S_{X,buf} (fp32, 64 columns; P overwrites its first 32) at (X*kSBuf + buf)*64; O_X (fp32, 64 columns) at kOCol + X*64.
// S is triple-buffered per tile: S(j+1), S(j+2) are computed while the softmax group still works on S(j), so the two groups drift apart
// instead of marching in lockstep (one in its exp2 phase while the other loads / stores TMEM and hands over to the MMA thread).
enum { BAR_Q_FULL = 0, BAR_Q_EMPTY = 1, BAR_K_FULL = 2, BAR_V_FULL = BAR_K_FULL + kStages, BAR_KV_EMPTY = BAR_V_FULL + kStages,
       BAR_S_FULL = BAR_KV_EMPTY + kStages /*[X][buf]*/, BAR_P_FULL = BAR_S_FULL + 2 * kSBuf /*[X][buf]*/, BAR_O_FULL = BAR_P_FULL + 2 * kSBuf /*[X][buf]*/, BAR_COUNT = BAR_O_FULL + 2 * kSBuf };
// S_FULL / P_FULL are per buffer: with two score tiles in flight a softmax group may run one tile ahead of the MMA thread, and a single
// barrier could then advance two phases before its consumer looks at it (parity waits cannot tell phase n from n + 2).

constexpr int kFwdSmem = 2 * kQBytes /*Q A,B*/ + 2 * kStages * kKVBytes /*K,V ring*/ + BAR_COUNT * 8 + 64 + 1024;

template <bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
mhsa_ws_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv, FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB alignment by pointer arithmetic: keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* sQ = smem;                              // [2][16 KB]
  uint8_t* sK = sQ + 2 * kQBytes;                  // [kStages][8 KB]
  uint8_t* sV = sK + kStages * kKVBytes;           // [kStages][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStages * kKVBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, dim = a.dim;
  const int nkv = (T + kKV - 1) / kKV;

  if (tid == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    for (int i = 0; i < BAR_COUNT; ++i) {
      int count = 1;
      if (i >= BAR_P_FULL && i < BAR_P_FULL + 2 * kSBuf) count = 4;                          // one arrival per softmax warp
      if (i == BAR_Q_EMPTY || (i >= BAR_KV_EMPTY && i < BAR_KV_EMPTY + kStages)) count = 2;   // one per MMA issuer
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
        mbar_arrive_expect_tx(&bars[BAR_Q_FULL], activeB ? 2 * kQBytes : kQBytes);
        tma_load_3d(sQ, &tma_q, &bars[BAR_Q_FULL], h * kD, q0, b);
        if (activeB) tma_load_3d(sQ + kQBytes, &tma_q, &bars[BAR_Q_FULL], h * kD, q0 + kTile, b);
        for (int j = 0; j < nkv; ++j, ++kv_iter) {
          const int st = kv_iter % kStages;
          mbar_wait(&bars[BAR_KV_EMPTY + st], ((kv_iter / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[BAR_K_FULL + st], kKVBytes);
          tma_load_3d(sK + st * kKVBytes, &tma_kv, &bars[BAR_K_FULL + st], dim + h * kD, j * kKV, b);
          mbar_arrive_expect_tx(&bars[BAR_V_FULL + st], kKVBytes);
          tma_load_3d(sV + st * kKVBytes, &tma_kv, &bars[BAR_V_FULL + st], 2 * dim + h * kD, j * kKV, b);
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: warp 1 drives tile A, warp 2 tile B =====================
      // A single issuing thread was the bottleneck of this kernel (~65 clk per tcgen05.mma + ~100 per barrier wait / commit, 16 MMAs of
      // only 32-48 clk each per K/V tile): two issuers halve that serial chain.  Control flow is warp-uniform (all lanes wait on the
      // barriers); only the tcgen05 instructions sit under elect_one, which keeps the compiler from wrapping each one in a divergence loop.
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      const int X = warp - 1;
      const uint32_t q_addr = smem_u32(sQ + X * kQBytes);
      const uint32_t tO = tmem + kOCol + X * 64;
      uint32_t kv_base = 0, work = 0;
      uint32_t g = 0;    // S tiles issued so far (buffer = g % kSBuf)
      uint32_t pc = 0;   // P tiles consumed so far
      auto issue_s = [&](uint32_t it) {   // S = Q K(it)^T  (128 x 64 x 64) into buffer g % kSBuf
        mbar_wait(&bars[BAR_K_FULL + it % kStages], (it / kStages) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + (it % kStages) * kKVBytes);
        const uint32_t buf = g % kSBuf;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + (X * kSBuf + buf) * 64, make_sw128_desc(q_addr + k * 32, 16, 1024), make_sw128_desc(k_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&bars[BAR_S_FULL + kSBuf * X + buf]);
        }
        __syncwarp();
        ++g;
      };
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++work, kv_base += nkv) {
        const int qp = item % a.pairs;
        const bool activeB = qp * 2 * kTile + kTile < T;
        const int releases = (X == 0 && !activeB) ? 2 : 1;
        // Both issuers observe EVERY Q_FULL phase, even for an item tile B sits out: a waiter that skipped a phase would be fooled by the
        // parity of the phase before it (and, running a whole item ahead, by the K/V ring barriers two wraps behind).
        mbar_wait(&bars[BAR_Q_FULL], work & 1);
        tc_fence_after();
        if (X == 1 && !activeB) continue;       // warp 1 then releases the shared stages for both
        for (int jj = 0; jj < kSBuf && jj < nkv; ++jj) issue_s(kv_base + jj);   // prologue: kSBuf score tiles ahead
        if (nkv <= kSBuf && elect_one())
          for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_Q_EMPTY]);
        for (int j = 0; j < nkv; ++j) {
          const uint32_t itj = kv_base + j;
          const int st = itj % kStages;
          mbar_wait(&bars[BAR_V_FULL + st], (itj / kStages) & 1);
          const uint32_t v_addr = smem_u32(sV + st * kKVBytes);
          const uint32_t buf = pc % kSBuf;
          mbar_wait(&bars[BAR_P_FULL + kSBuf * X + buf], (pc / kSBuf) & 1);
          tc_fence_after();
          if (elect_one()) {                      // O (+)= P V  (128 x 64 x 64), P read from TMEM; frees S buffer j % kSBuf (in-order execution)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(tO, tmem + (X * kSBuf + buf) * 64 + 8 * k, make_sw128_desc(v_addr + k * a.mn_kadv, a.mn_lbo, a.mn_sbo), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bars[BAR_O_FULL + kSBuf * X + buf]);
          }
          __syncwarp();
          ++pc;
          if (j + kSBuf < nkv) issue_s(itj + kSBuf);   // ... which S(j + kSBuf) then overwrites
          if (elect_one()) {
            if (j + kSBuf + 1 == nkv)                  // the last S MMAs (tile nkv-1) were just issued: Q is free when they complete
              for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_Q_EMPTY]);
            for (int rr = 0; rr < releases; ++rr) umma_commit(&bars[BAR_KV_EMPTY + st]);   // K(j) was read by S(j), V(j) by PV(j)
          }
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
    const uint32_t tO = tmem + kOCol + X * 64 + lane_off;   // O (fp32, 64 columns)
    const float c2 = a.scale * kLog2e;
    uint32_t g = 0;                           // score tiles consumed so far by this group
    const MhsaDrop drop = kDrop ? mhsa_salted(a.drop) : a.drop;
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
      const int bh = item / a.pairs, qp = item - bh * a.pairs;
      const int h = bh % a.H, b = bh / a.H;
      const int q0 = qp * 2 * kTile + X * kTile;
      if (q0 >= T) continue;                  // tile B of the last pair may be empty (never for tile A)
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < nkv; ++j, ++g) {
        const uint32_t buf = g % kSBuf;
        const uint32_t tS = tmem + (X * kSBuf + buf) * 64 + lane_off;
        mbar_wait(&bars[BAR_S_FULL + kSBuf * X + buf], (g / kSBuf) & 1);
        tc_fence_after();
        if ((a.dbg & 3) == 3) {   // timing experiment: the MMA / TMA pipeline alone
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_P_FULL + kSBuf * X + buf]);
          continue;
        }
        float s[64];
        if ((a.dbg & 3) != 2) {
          tmem_ld_32x32(tS + 0, *reinterpret_cast<float(*)[32]>(&s[0]));
          tmem_ld_32x32(tS + 32, *reinterpret_cast<float(*)[32]>(&s[32]));
          tc_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) s[i] = 0.01f * (float)((i * 7 + r + j) & 63);
        }
        const int valid = T - j * kKV;        // key columns >= valid are zero-filled padding
        if (valid < kKV) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= valid) s[i] = -INFINITY;
        }
        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int i = 4; i < 64; i += 4) {
          mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]); mx2 = fmaxf(mx2, s[i + 2]); mx3 = fmaxf(mx3, s[i + 3]);
        }
        const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
        const bool resc = (m_new - m_used) * c2 > kRescaleThreshold;   // also true on the first tile (m_used = -inf)
        float alpha = 1.f;
        if (resc) {
          alpha = fast_ex2((m_used - m_new) * c2);
          m_used = m_new;
        }
        const float mc = m_used * c2;
        const float2 c2v = make_float2(c2, c2), mcv = make_float2(-mc, -mc);
        float2 rs01 = make_float2(0.f, 0.f), rs23 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {           // 32 scores -> 16 packed registers -> P columns 16c .. 16c+15 (over S columns already read)
          uint32_t pk[16];
          uint32_t keep = 0xFFFFFFFFu;          // dropout decisions of these 32 keys (the row sum l stays that of the un-dropped softmax)
          if (kDrop) keep = mhsa_keep16(drop, bh, q0 + r, 4 * j + 2 * c) | (mhsa_keep16(drop, bh, q0 + r, 4 * j + 2 * c + 1) << 16);
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float2 pa = ffma2(make_float2(s[32 * c + 2 * i], s[32 * c + 2 * i + 1]), c2v, mcv);
            float2 pb = ffma2(make_float2(s[32 * c + 2 * i + 2], s[32 * c + 2 * i + 3]), c2v, mcv);
            if ((a.dbg & 3) != 1) { pa.x = fast_ex2(pa.x); pa.y = fast_ex2(pa.y); pb.x = fast_ex2(pb.x); pb.y = fast_ex2(pb.y); }
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
          tmem_st_32x16(tS + 16 * c, pk);
        }
        const float rs0 = rs01.x, rs1 = rs01.y, rs2 = rs23.x, rs3 = rs23.y;
        l = fmaf(l, alpha, (rs0 + rs1) + (rs2 + rs3));
        if (j > 0 && __any_sync(0xffffffffu, resc)) {
          // O (accumulated by PV of the previous iterations) must be in the units of the new maximum before PV(j) adds to it
          // PV(g-1) signals the barrier of ITS score buffer: a group can be kSBuf-1 tiles ahead of the PVs, so one barrier per tile would
          // alias phases; with one per buffer, S(g) being complete implies PV(g-kSBuf) is, i.e. the barrier is at most one phase behind
          mbar_wait(&bars[BAR_O_FULL + kSBuf * X + (g - 1) % kSBuf], ((g - 1) / kSBuf) & 1);
          tc_fence_after();
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
        if (lane == 0) mbar_arrive(&bars[BAR_P_FULL + kSBuf * X + buf]);
      }
      // ---- epilogue: O / l -> bf16, lse
      mbar_wait(&bars[BAR_O_FULL + kSBuf * X + (g - 1) % kSBuf], ((g - 1) / kSBuf) & 1);
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
      // the next item's PV(0) (accumulate = 0) is gated by this group's next P_FULL arrival, i.e. after these O reads: no extra barrier
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}

}  // namespace ws

int mhsa_ws_fwd(const gvk_mhsa_fwd_params* p, cudaStream_t stream) {
  using namespace ws;
  static bool configured = false;
  if (!configured) {
    int st = cuda_status(cudaFuncSetAttribute(mhsa_ws_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem), "mhsa_ws_fwd smem");
    if (st != GVK_OK) return st;
    st = cuda_status(cudaFuncSetAttribute(mhsa_ws_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem), "mhsa_ws_fwd (dropout) smem");
    if (st != GVK_OK) return st;
    configured = true;
  }
  const int dim = p->H * kD;
  CUtensorMap tq, tkv;
  int st = make_tma_3d_bf16(&tq, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, kTile, kD);
  if (st != GVK_OK) return st;
  st = make_tma_3d_bf16(&tkv, p->qkv, p->B, p->T, 3 * dim, p->ld, (uint64_t)p->T * p->ld, kKV, kD);
  if (st != GVK_OK) return st;
  FwdArgs a;
  a.B = p->B; a.T = p->T; a.H = p->H; a.dim = dim; a.scale = p->scale;
  a.out = reinterpret_cast<__nv_bfloat16*>(p->out);
  a.ld_out = p->ld_out;
  a.lse = p->lse;
  a.pairs = ((p->T + kTile - 1) / kTile + 1) / 2;
  a.num_items = p->B * p->H * a.pairs;
  a.mn_lbo = 8192; a.mn_sbo = 1024; a.mn_kadv = 2048;
  { const char* e = getenv("GVK_WS_DBG"); a.dbg = e ? atoi(e) : 0; }
  const int grid = std::min(a.num_items, sm_count());
  a.drop = make_mhsa_drop(p->drop_p, p->seed, p->seed_salt);
  if (p->drop_p > 0.f)
    mhsa_ws_fwd_kernel<true><<<grid, kThreads, kFwdSmem, stream>>>(tq, tkv, a);
  else
    mhsa_ws_fwd_kernel<false><<<grid, kThreads, kFwdSmem, stream>>>(tq, tkv, a);
  GVK_CHECK_LAUNCH("mhsa_ws_fwd");
  return GVK_OK;
}

}  // namespace gvk
