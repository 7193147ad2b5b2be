// gvk_gemm.cu — dense TN GEMM with fused epilogue (see include/gvk.h: gvk_gemm).
//
//  * bf16 path: persistent, warp-specialised tcgen05 kernel.  TMA (SWIZZLE_128B) feeds a 4..6-stage smem ring, one elected
//    thread issues tcgen05.mma (UMMA 128 x BN x 16, fp32 accumulators in TMEM, double-buffered so the epilogue of tile i
//    overlaps the main loop of tile i+1), eight epilogue warps drain TMEM with tcgen05.ld, transpose through a private
//    smem patch so that every global access of the fused epilogue is row-contiguous.
//  * fp32 path: exact FFMA kernel with the same epilogue (the "fp32 mode" used for 1e-4 parity).
//
// Replaces the nn.Linear / Conv3d-as-GEMM calls of the reference (model/vision_transformer.py:31-35,53-58,
// model/gaviko.py:383-385) and their dgrad counterparts in autograd.
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {

struct EpiArgs {
  const float* bias;
  const float* ssf_scale;
  const float* ssf_shift;
  int act;
  void* aux;
  int aux_dtype;
  int ld_aux;
  const float* pos;
  int rows_per_batch;
  int out_batch_rows;
  int out_row_offset;
  const float* res1;
  int ld_res1;
  const float* res2;
  int ld_res2;
  void* out;
  int out_dtype;
  int ld_out;
  float* out2;
  int ld_out2;
  int N;
};

// Per-column constants of the epilogue, loaded once per (lane, column).
struct EpiCol {
  float bias, sc, sh;
};
__device__ __forceinline__ EpiCol epi_col(const EpiArgs& e, int n, bool valid) {
  EpiCol c;
  c.bias = (e.bias && valid) ? e.bias[n] : 0.f;
  c.sc = (e.ssf_scale && valid) ? e.ssf_scale[n] : 1.f;
  c.sh = (e.ssf_shift && valid) ? e.ssf_shift[n] : 0.f;
  return c;
}
__device__ __forceinline__ void epi_elem(const EpiArgs& e, int m, int n, float v, const EpiCol& c) {
  v += c.bias;
  if (e.ssf_scale) v = v * c.sc + c.sh;
  if (e.act == GVK_ACT_GELU) {
    if (e.aux) st_dyn(e.aux, (size_t)m * e.ld_aux + n, e.aux_dtype, v);
    v = gelu_erf(v);
  } else if (e.act == GVK_ACT_GELU_BWD) {
    v *= gelu_erf_grad(ld_dyn(e.aux, (size_t)m * e.ld_aux + n, e.aux_dtype));
  } else if (e.act == GVK_ACT_GELU_SAVE_GRAD) {
    if (e.aux) st_dyn(e.aux, (size_t)m * e.ld_aux + n, e.aux_dtype, gelu_erf_grad(v));
    v = gelu_erf(v);
  } else if (e.act == GVK_ACT_MUL_AUX) {
    v *= ld_dyn(e.aux, (size_t)m * e.ld_aux + n, e.aux_dtype);
  }
  int orow = m;
  if (e.rows_per_batch > 0) {
    const int b = m / e.rows_per_batch;
    const int r = m - b * e.rows_per_batch;
    if (e.pos) v += e.pos[(size_t)r * e.N + n];
    orow = b * e.out_batch_rows + e.out_row_offset + r;
  }
  if (e.res1) v += e.res1[(size_t)m * e.ld_res1 + n];
  if (e.res2) v += e.res2[(size_t)m * e.ld_res2 + n];
  st_dyn(e.out, (size_t)orow * e.ld_out + n, e.out_dtype, v);
  if (e.out2) e.out2[(size_t)m * e.ld_out2 + n] = v;
}


// -------------------------------------------------------------------------------------------------
// Fast epilogues (compile-time specialised, 128-bit accesses).  One warp owns a 32-row x 32-column patch:
// thread = row after tcgen05.ld, then an XOR-swizzled smem transpose makes lane = (sub-row, 4-column vector) so that every
// global access is a contiguous 128 B (fp32) / 64 B (bf16) row segment.  Loads the epilogue depends on (residual, saved GELU
// pre-activation) are issued before the TMEM wait so their latency overlaps it.
// -------------------------------------------------------------------------------------------------
enum { EPI_GENERIC = 0, EPI_STORE_BF16 = 1, EPI_BIAS_GELU_BF16 = 2, EPI_BIAS_RES_F32 = 3, EPI_GELU_BWD_BF16 = 4, EPI_STORE_F32 = 5, EPI_BIAS_GELU_SAVEGRAD_BF16 = 6, EPI_MUL_AUX_BF16 = 7,
       EPI_BIAS_BF16 = 8, EPI_BIAS_F32 = 9 };   // 8 / 9: bias only (SSF sites folded into the weights, Linear outputs that feed a separate dropout)
template <int EPI>
constexpr bool kEpiBias = (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RES_F32 || EPI == EPI_BIAS_GELU_SAVEGRAD_BF16 || EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_F32);

// Standard-normal CDF Phi(x) = 0.5 (1 + erf(x / sqrt 2)) with |abs err| < 2e-7 (Abramowitz-Stegun 7.1.26), branch-free:
// rcp.approx + ex2.approx + 7 fma/mul + select.  Also returns e = exp(-x^2 / 2) for the GELU derivative.  Used only on the bf16
// path, whose results are rounded to bf16 anyway (the exact-fp32 kernel keeps erff).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float norm_cdf_fast(float x, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  e = ex2_approx(x * x * -0.72134752044448170f);  // exp(-x^2/2)
  const float q = p * t * e;                      // 0.5 * erfc(|x| / sqrt 2)
  return x < 0.f ? q : 1.0f - q;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float e;
  return x * norm_cdf_fast(x, e);
}
// GELU and its derivative together: they share the CDF and exp(-x^2/2)
__device__ __forceinline__ float gelu_and_grad_fast(float x, float& grad) {
  float e;
  const float cdf = norm_cdf_fast(x, e);
  grad = fmaf(x * 0.39894228040143268f, e, cdf);
  return x * cdf;
}
// Two elements at a time on the packed fp32 instructions (FFMA2 / FMUL2 / FADD2): the epilogue of the fc1 forward GEMM is bound by its own
// instruction count (8 warps against a 5 us main loop), and the polynomial / products of the pair share issue slots this way.
__device__ __forceinline__ void gelu_and_grad_fast2(float2 x, float2& y, float2& grad) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  float2 t = ffma2(ax, make_float2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f), make_float2(1.0f, 1.0f));
  t.x = rcp_approx(t.x);
  t.y = rcp_approx(t.y);
  float2 p = ffma2(make_float2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t, make_float2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
  p = ffma2(p, t, make_float2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
  p = ffma2(p, t, make_float2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
  p = ffma2(p, t, make_float2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
  float2 e = fmul2(fmul2(x, x), make_float2(-0.72134752044448170f, -0.72134752044448170f));
  e.x = ex2_approx(e.x);
  e.y = ex2_approx(e.y);                                   // exp(-x^2 / 2)
  const float2 q = fmul2(fmul2(p, t), e);                  // 0.5 erfc(|x| / sqrt 2)
  const float2 cdf = make_float2(x.x < 0.f ? q.x : 1.0f - q.x, x.y < 0.f ? q.y : 1.0f - q.y);
  y = fmul2(x, cdf);
  grad = ffma2(fmul2(x, make_float2(0.39894228040143268f, 0.39894228040143268f)), e, cdf);
}
// GELU and its derivative through ONE SFU op per element: Phi(x) ~ 0.5 (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) (|error| < 5e-4, below the
// bf16 rounding of the results it feeds; tanh.approx adds 2^-11), derivative of that form taken analytically — 12 packed instructions per
// PAIR of elements against 21 for the erfc polynomial above.  The fc1-forward epilogue is bound by the SM's issue slots (26 k warp
// instructions per 128 x 256 tile against a 6.1 k clk main loop), so the instruction count is what sets its speed.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_and_grad_tanh2(float2 x, float2& y, float2& grad) {
  constexpr float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float2 x2 = fmul2(x, x);
  const float2 inner = ffma2(x2, make_float2(k0 * k1, k0 * k1), make_float2(k0, k0));          // k0 (1 + k1 x^2)
  float2 t = fmul2(x, inner);
  t.x = tanh_approx(t.x);
  t.y = tanh_approx(t.y);
  const float2 cdf = ffma2(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
  y = fmul2(x, cdf);
  const float2 sech2 = ffma2(make_float2(-t.x, -t.y), t, make_float2(1.0f, 1.0f));
  const float2 dinner = ffma2(x2, make_float2(3.0f * k0 * k1, 3.0f * k0 * k1), make_float2(k0, k0));   // d/dx [k0 (x + k1 x^3)]
  grad = ffma2(fmul2(fmul2(x, make_float2(0.5f, 0.5f)), sech2), dinner, cdf);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  constexpr float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return x * fmaf(tanh_approx(x * fmaf(x * x, k0 * k1, k0)), 0.5f, 0.5f);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float e;
  const float cdf = norm_cdf_fast(x, e);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 u) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 float4_to_bf16x4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  return u;
}

// The saved-activation operand of the GELU-backward / multiply-by-aux epilogues streams from HBM.  A warp requests the operand of all its
// patches of a tile (8 KB) before it waits for the tile's accumulator, so 64 KB per SM are in flight and the latency hides behind the main
// loop; loading patch by patch put four DRAM round trips (~8 us) on the critical path of a 5 us tile.  Rows past M are clamped, not
// predicated: a select on the loaded value would make the warp wait for each load where it is issued.
template <int EPI>
constexpr bool kEpiAuxIn = (EPI == EPI_GELU_BWD_BF16 || EPI == EPI_MUL_AUX_BF16);
template <int EPI>
__device__ __forceinline__ void epilogue_prefetch_aux(const EpiArgs& e, int row0, int col0, int M, int lane, uint2 (&aux)[8]) {
  if constexpr (kEpiAuxIn<EPI>) {
    if (col0 < e.N && row0 < M) {   // warp-uniform
      const int sr = lane >> 3, col = col0 + (lane & 7) * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int m = row0 + it * 4 + sr;
        aux[it] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(e.aux) + (size_t)min(m, M - 1) * e.ld_aux + col);
      }
    }
  }
}

// The fp32 residual of a patch (32 registers) is requested one patch ahead of its use.
template <int EPI>
__device__ __forceinline__ void epilogue_prefetch_res(const EpiArgs& e, int row0, int col0, int M, int lane, float4 (&res)[8]) {
  if constexpr (EPI == EPI_BIAS_RES_F32) {
    if (col0 < e.N && row0 < M) {   // warp-uniform
      const int sr = lane >> 3, col = col0 + (lane & 7) * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) res[it] = *reinterpret_cast<const float4*>(e.res1 + (size_t)min(row0 + it * 4 + sr, M - 1) * e.ld_res1 + col);
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_patch_fast(const EpiArgs& e, uint32_t taddr, float* patch, int row0, int col0, int M, int lane, const uint2 (&aux_in)[8],
                                                    const float4 (&res)[8]) {
  const int sr = lane >> 3, cv = lane & 7;  // sub-row 0..3, 4-column vector 0..7
  const int col = col0 + cv * 4;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (kEpiBias<EPI>) bias = *reinterpret_cast<const float4*>(e.bias + col);
  // ---- TMEM -> registers (thread = row) -> swizzled smem  (loading the next patch's accumulator during this one's math was tried: no gain)
  float v[32];
  tmem_ld_32x32(taddr, v);
  tc_wait_ld();
  {
    float4* prow = reinterpret_cast<float4*>(patch) + lane * 8;
#pragma unroll
    for (int c = 0; c < 8; ++c) prow[c ^ (lane & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  }
  __syncwarp();
  // ---- lane = (sub-row, vector): fused elementwise + coalesced stores
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + sr;
    const int m = row0 + r;
    float4 x = reinterpret_cast<const float4*>(patch)[r * 8 + (cv ^ (r & 7))];
    const bool live = m < M;   // rows past M are computed like any other and only their stores are predicated: a `continue` here put a
                               // branch between the eight unrolled iterations and kept the scheduler from interleaving their math
    if constexpr (kEpiBias<EPI>) {
      x.x += bias.x; x.y += bias.y; x.z += bias.z; x.w += bias.w;
    }
    if constexpr (EPI == EPI_BIAS_GELU_SAVEGRAD_BF16) {
      float4 gr;
      float2 y01, y23, g01, g23;
      gelu_and_grad_tanh2(make_float2(x.x, x.y), y01, g01);
      gelu_and_grad_tanh2(make_float2(x.z, x.w), y23, g23);
      x = make_float4(y01.x, y01.y, y23.x, y23.y);
      gr = make_float4(g01.x, g01.y, g23.x, g23.y);
      if (e.aux && live) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.aux) + (size_t)m * e.ld_aux + col) = float4_to_bf16x4(gr);
    }
    if constexpr (EPI == EPI_MUL_AUX_BF16) {
      const float4 a4 = bf16x4_to_float4(aux_in[it]);
      x.x *= a4.x; x.y *= a4.y; x.z *= a4.z; x.w *= a4.w;
    }
    if constexpr (EPI == EPI_BIAS_GELU_BF16) {
      if (e.aux && live) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.aux) + (size_t)m * e.ld_aux + col) = float4_to_bf16x4(x);
      x.x = gelu_tanh(x.x); x.y = gelu_tanh(x.y); x.z = gelu_tanh(x.z); x.w = gelu_tanh(x.w);
    }
    if constexpr (EPI == EPI_GELU_BWD_BF16) {
      const float4 pre = bf16x4_to_float4(aux_in[it]);
      x.x *= gelu_grad_fast(pre.x); x.y *= gelu_grad_fast(pre.y); x.z *= gelu_grad_fast(pre.z); x.w *= gelu_grad_fast(pre.w);
    }
    if constexpr (EPI == EPI_BIAS_RES_F32) {
      x.x += res[it].x; x.y += res[it].y; x.z += res[it].z; x.w += res[it].w;
    }
    if (live) {
      if constexpr (EPI == EPI_BIAS_RES_F32 || EPI == EPI_STORE_F32 || EPI == EPI_BIAS_F32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (size_t)m * e.ld_out + col) = x;
      } else {
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.out) + (size_t)m * e.ld_out + col) = float4_to_bf16x4(x);
      }
    }
  }
  __syncwarp();
}

// =================================================================================================
// tcgen05 kernel
// =================================================================================================
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 128 + kEpiWarps * 32;
constexpr int kEpiPatch = 32 * 33;  // floats per epilogue warp

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kRing = STAGES * (kABytes + kBBytes);
  static constexpr int kEpi = kEpiWarps * kEpiPatch * 4;
  static constexpr int kBars = (2 * STAGES + 4) * 8 + 16;
  static constexpr int kTotal = kRing + kEpi + kBars + 1024;  // +1024: manual alignment slack
};

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_sm100_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, EpiArgs e, int M, int N, int K) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB alignment by pointer arithmetic: keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::kABytes;
  float* epi_patch = reinterpret_cast<float*>(smem + L::kRing);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kRing + L::kEpi);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512: power of two

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + kBM - 1) / kBM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / kBK;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t / num_n, n_blk = t % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], L::kABytes + L::kBBytes);
          tma_load_2d(sA + stage * L::kABytes, &tma_a, &full_bar[stage], kb * kBK, m_blk * kBM);
          tma_load_2d(sB + stage * L::kBBytes, &tma_b, &full_bar[stage], kb * kBK, n_blk * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int lt = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
        const int as = lt & 1;
        const uint32_t aphase = (lt >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * L::kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * L::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = make_sw128_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = make_sw128_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once the MMAs above have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> regs -> private smem transpose -> fused elementwise -> global =====
    const int ew = warp - 4;
    const int q = warp & 3;   // TMEM lane group this warp may access
    const int h = ew >> 2;    // column half
    float* patch = epi_patch + ew * kEpiPatch;
    int lt = 0;
    constexpr int kPatches = BN / 64;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      const int m_blk = t / num_n, n_blk = t % num_n;
      const int as = lt & 1;
      const uint32_t aphase = (lt >> 1) & 1;
      const int row0 = m_blk * kBM + q * 32;
      constexpr bool kRes = EPI == EPI_BIAS_RES_F32;
      uint2 aux_in[kEpiAuxIn<EPI> ? kPatches : 1][8];
      float4 res[kRes ? 2 : 1][8];
      const int tile_col0 = n_blk * BN + h * (BN / 2);
      if constexpr (kEpiAuxIn<EPI>) {
#pragma unroll
        for (int c = 0; c < kPatches; ++c) epilogue_prefetch_aux<EPI>(e, row0, tile_col0 + c * 32, M, lane, aux_in[c]);
      }
      epilogue_prefetch_res<EPI>(e, row0, tile_col0, M, lane, res[0]);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + h * (BN / 2);
#pragma unroll((kEpiAuxIn<EPI> || kRes) ? kPatches : 1)
      for (int c = 0; c < kPatches; ++c) {
        const int colt = h * (BN / 2) + c * 32;
        const int col0 = n_blk * BN + colt;
        if (c + 1 < kPatches) epilogue_prefetch_res<EPI>(e, row0, col0 + 32, M, lane, res[kRes ? (c + 1) & 1 : 0]);
        if (col0 < N && row0 < M) {  // warp-uniform
          const uint32_t taddr = taddr0 + c * 32;
          if constexpr (EPI != EPI_GENERIC) {
            epilogue_patch_fast<EPI>(e, taddr, patch, row0, col0, M, lane, aux_in[kEpiAuxIn<EPI> ? c : 0], res[kRes ? c & 1 : 0]);
          } else {
            float v[32];
            tmem_ld_32x32(taddr, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) patch[lane * 33 + j] = v[j];
            __syncwarp();
            const int n = col0 + lane;
            const bool nvalid = n < N;
            const EpiCol ec = epi_col(e, n, nvalid);
            const int rmax = min(32, M - row0);
            if (nvalid) {
#pragma unroll 4
              for (int r = 0; r < rmax; ++r) epi_elem(e, row0 + r, n, patch[r * 33 + lane], ec);
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// exact fp32 FFMA kernel (64x64x16 tiles, 4x4 micro-tiles)
// =================================================================================================
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, int lda, int ldb, EpiArgs e,
                                                         int M, int N, int K) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int lr = tid >> 2;         // 0..63 row inside tile
  const int lk = (tid & 3) * 4;    // 0,4,8,12
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (m0 + lr < M && k0 + lk < K) av = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lr) * lda + k0 + lk);
    if (n0 + lr < N && k0 + lk < K) bv = *reinterpret_cast<const float4*>(B + (size_t)(n0 + lr) * ldb + k0 + lk);
    sA[lk + 0][lr] = av.x; sA[lk + 1][lr] = av.y; sA[lk + 2][lr] = av.z; sA[lk + 3][lr] = av.w;
    sB[lk + 0][lr] = bv.x; sB[lk + 1][lr] = bv.y; sB[lk + 2][lr] = bv.z; sB[lk + 3][lr] = bv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx + 16 * j;
    if (n >= N) continue;
    const EpiCol ec = epi_col(e, n, true);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < M) epi_elem(e, m, n, acc[i][j], ec);
    }
  }
}

// =================================================================================================
// host
// =================================================================================================
static EpiArgs to_epi(const gvk_gemm_params* p) {
  EpiArgs e;
  e.bias = p->bias; e.ssf_scale = p->ssf_scale; e.ssf_shift = p->ssf_shift; e.act = p->act;
  e.aux = p->aux; e.aux_dtype = p->aux_dtype; e.ld_aux = p->ld_aux;
  e.pos = p->pos; e.rows_per_batch = p->rows_per_batch; e.out_batch_rows = p->out_batch_rows; e.out_row_offset = p->out_row_offset;
  e.res1 = p->res1; e.ld_res1 = p->ld_res1; e.res2 = p->res2; e.ld_res2 = p->ld_res2;
  e.out = p->out; e.out_dtype = p->out_dtype; e.ld_out = p->ld_out; e.out2 = p->out2; e.ld_out2 = p->ld_out2;
  e.N = p->N;
  return e;
}

template <int BN, int STAGES, int EPI>
static int launch_bf16(const gvk_gemm_params* p, const EpiArgs& e, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  static bool configured = false;
  auto kern = gemm_bf16_sm100_kernel<BN, STAGES, EPI>;
  if (!configured) {
    int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal), "gemm smem attribute");
    if (st != GVK_OK) return st;
    configured = true;
  }
  CUtensorMap ta, tb;
  int st = make_tma_2d_bf16(&ta, p->a, p->M, p->K, p->lda, kBM, kBK);
  if (st != GVK_OK) return st;
  st = make_tma_2d_bf16(&tb, p->b, p->N, p->K, p->ldb, BN, kBK);
  if (st != GVK_OK) return st;
  const int tiles = ((p->M + kBM - 1) / kBM) * ((p->N + BN - 1) / BN);
  const int grid = std::min(tiles, sm_count());
  kern<<<grid, kGemmThreads, L::kTotal, stream>>>(ta, tb, e, p->M, p->N, p->K);
  GVK_CHECK_LAUNCH("gemm_bf16_sm100");
  return GVK_OK;
}

int gemm_dispatch(const gvk_gemm_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->a && p->b && p->out, "gvk_gemm: null operand");
  GVK_CHECK_ARG(p->M > 0 && p->N > 0 && p->K > 0, "gvk_gemm: non-positive shape M=%d N=%d K=%d", p->M, p->N, p->K);
  GVK_CHECK_ARG((p->act != GVK_ACT_GELU_BWD && p->act != GVK_ACT_MUL_AUX) || p->aux, "gvk_gemm: GELU_BWD / MUL_AUX need aux");
  GVK_CHECK_ARG(p->rows_per_batch >= 0, "gvk_gemm: rows_per_batch < 0");
  const EpiArgs e = to_epi(p);
  if (p->ab_dtype == GVK_BF16) {
    GVK_CHECK_ARG(p->K % kBK == 0, "gvk_gemm(bf16): K=%d must be a multiple of %d", p->K, kBK);
    GVK_CHECK_ARG(p->lda % 8 == 0 && p->ldb % 8 == 0, "gvk_gemm(bf16): lda/ldb must be multiples of 8");
    GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(p->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->b) & 15) == 0,
                  "gvk_gemm(bf16): operands must be 16-byte aligned");
    // pick a compile-time specialised epilogue when the request matches one of the hot-path shapes
    int epi = EPI_GENERIC;
    const bool plain = !p->ssf_scale && !p->ssf_shift && !p->pos && p->rows_per_batch == 0 && !p->res2 && !p->out2 && p->N % 32 == 0 && p->ld_out % 4 == 0 &&
                       (reinterpret_cast<uintptr_t>(p->out) & 15) == 0;
    if (plain) {
      const bool aux_ok = !p->aux || (p->aux_dtype == GVK_BF16 && p->ld_aux % 4 == 0 && (reinterpret_cast<uintptr_t>(p->aux) & 7) == 0);
      const bool res_ok = p->res1 && p->ld_res1 % 4 == 0 && (reinterpret_cast<uintptr_t>(p->res1) & 15) == 0;
      const bool bias_ok = p->bias && (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0;
      if (!p->bias && p->act == GVK_ACT_NONE && !p->res1 && !p->aux) epi = p->out_dtype == GVK_BF16 ? EPI_STORE_BF16 : EPI_STORE_F32;
      else if (bias_ok && p->act == GVK_ACT_GELU && !p->res1 && p->out_dtype == GVK_BF16 && aux_ok) epi = EPI_BIAS_GELU_BF16;
      else if (bias_ok && p->act == GVK_ACT_NONE && res_ok && p->out_dtype == GVK_F32 && !p->aux) epi = EPI_BIAS_RES_F32;
      else if (!p->bias && p->act == GVK_ACT_GELU_BWD && !p->res1 && p->out_dtype == GVK_BF16 && p->aux && aux_ok) epi = EPI_GELU_BWD_BF16;
      else if (bias_ok && p->act == GVK_ACT_GELU_SAVE_GRAD && !p->res1 && p->out_dtype == GVK_BF16 && aux_ok) epi = EPI_BIAS_GELU_SAVEGRAD_BF16;
      else if (!p->bias && p->act == GVK_ACT_MUL_AUX && !p->res1 && p->out_dtype == GVK_BF16 && p->aux && aux_ok) epi = EPI_MUL_AUX_BF16;
      else if (bias_ok && p->act == GVK_ACT_NONE && !p->res1 && !p->aux) epi = p->out_dtype == GVK_BF16 ? EPI_BIAS_BF16 : EPI_BIAS_F32;
    }
#define GVK_GEMM_CASE(E)                                               \
  case E:                                                              \
    if (p->N % 256 == 0) return launch_bf16<256, 4, E>(p, e, stream); \
    return launch_bf16<128, 6, E>(p, e, stream);
    switch (epi) {
      GVK_GEMM_CASE(EPI_STORE_BF16)
      GVK_GEMM_CASE(EPI_BIAS_GELU_BF16)
      GVK_GEMM_CASE(EPI_BIAS_RES_F32)
      GVK_GEMM_CASE(EPI_GELU_BWD_BF16)
      GVK_GEMM_CASE(EPI_STORE_F32)
      GVK_GEMM_CASE(EPI_BIAS_GELU_SAVEGRAD_BF16)
      GVK_GEMM_CASE(EPI_MUL_AUX_BF16)
      GVK_GEMM_CASE(EPI_BIAS_BF16)
      GVK_GEMM_CASE(EPI_BIAS_F32)
      default:
        GVK_GEMM_CASE(EPI_GENERIC)
    }
#undef GVK_GEMM_CASE
  }
  if (p->ab_dtype == GVK_F32) {
    GVK_CHECK_ARG(p->K % 4 == 0 && p->lda % 4 == 0 && p->ldb % 4 == 0, "gvk_gemm(f32): K, lda, ldb must be multiples of 4");
    dim3 grid((p->N + 63) / 64, (p->M + 63) / 64);
    gemm_f32_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(p->a), reinterpret_cast<const float*>(p->b), p->lda, p->ldb, e,
                                              p->M, p->N, p->K);
    GVK_CHECK_LAUNCH("gemm_f32");
    return GVK_OK;
  }
  set_last_error("gvk_gemm: unsupported ab_dtype %d", p->ab_dtype);
  return GVK_ERR_UNSUPPORTED;
}

}  // namespace gvk
