// gvk_evp.cu — the two primitives --method evp (reference src/model/evp.py) needs beyond the frozen-ViT kernel set:
//
//  (1) gvk_wgrad: a general weight-gradient product  dw[i, j] += sum_m a[row_a(m), i] * b[row_b(m), j]  (both operands "MN-major": the
//      reduction runs over ROWS).  EVP's prompt generator has rank dim / scale_factor (192 at ViT-B with the shipped evp.yaml), far past
//      the r <= 32 limit of the rank-r kernels: autograd's Linear / Conv3d weight gradients of model/evp.py:42-52,85-95.
//      GVK_PREC_TF32: mma.sync m16n8k8 (tf32 operands rounded to nearest, fp32 accumulate) on 128 x 64 output tiles, rows split over
//      CTAs, fp32 atomics into dw.  GVK_PREC_FP32: the same tiling with exact FFMAs (the 1e-4 parity mode).
//
//  (2) gvk_hfreq_filter: PromptGenerator.fft (model/evp.py:124-146) in closed form.  For a 5-D volume the reference's fft2 / all-axes
//      fftshift / 4-index mask assignment amounts to: on the depth slices selected by `hit`, remove a band of H-frequencies (every
//      W-frequency is kept), take the real part and the absolute value; other slices only take |x|.  Removing a frequency band along H
//      is a real H x H matrix F applied to the columns of the slice (F = I - Re(IDFT diag(cut) DFT), built by the host in fp64), so
//      the kernel is a batched (H x H) @ (H x W) product with |.| in the epilogue — no FFT.  Exact fp32 FMAs.
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {

// ------------------------------------------------------------------------------------------------
// (1) general weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int kWgTA = 128;     // tile extent along a's columns (rows of dw)
constexpr int kWgTB = 64;      // tile extent along b's columns (columns of dw)
constexpr int kWgRowsStage = 32;
constexpr int kWgThreads = 256;
constexpr int kWgLdA = kWgTA + 8;   // (t * ld + g) % 32 distinct over a warp for ld % 32 == 8: conflict-free fragment loads
constexpr int kWgLdB = kWgTB + 8;

template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ size_t wg_row(int m, int rpb, int batch_rows) {
  if (rpb <= 0) return (size_t)m;
  const int b = m / rpb;
  return (size_t)b * batch_rows + (m - b * rpb);
}

template <typename TA, typename TB, bool TF32>
__global__ void __launch_bounds__(kWgThreads) wgrad_kernel(gvk_wgrad_params p, int rows_per_cta) {
  __shared__ __align__(16) float sA[kWgRowsStage * kWgLdA];
  __shared__ __align__(16) float sB[kWgRowsStage * kWgLdB];
  const TA* A = reinterpret_cast<const TA*>(p.a);
  const TB* Bm = reinterpret_cast<const TB*>(p.b);
  const int i0 = blockIdx.x * kWgTA, j0 = blockIdx.y * kWgTB;
  const int m_begin = blockIdx.z * rows_per_cta, m_end = min(p.M, m_begin + rows_per_cta);
  const int tid = threadIdx.x;
  // staging map: A stage = 32 rows x 32 float4 -> 4 per thread; B stage = 32 rows x 16 float4 -> 2 per thread
  float4 ra[4], rb[2];
  auto load_stage = [&](int m0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * kWgThreads, row = idx >> 5, c4 = (idx & 31) * 4;
      const int m = m0 + row, col = i0 + c4;
      ra[q] = (m < m_end && col < p.na) ? ld4<TA>(A + wg_row(m, p.a_rows_per_batch, p.a_batch_rows) * p.lda + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kWgThreads, row = idx >> 4, c4 = (idx & 15) * 4;
      const int m = m0 + row, col = j0 + c4;
      rb[q] = (m < m_end && col < p.nb) ? ld4<TB>(Bm + wg_row(m, p.b_rows_per_batch, p.b_batch_rows) * p.ldb + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto cvt = [](float4 v) {
    if constexpr (TF32) return make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
    else return v;
  };
  auto store_stage = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * kWgThreads, row = idx >> 5, c4 = (idx & 31) * 4;
      *reinterpret_cast<float4*>(sA + row * kWgLdA + c4) = cvt(ra[q]);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kWgThreads, row = idx >> 4, c4 = (idx & 15) * 4;
      *reinterpret_cast<float4*>(sB + row * kWgLdB + c4) = cvt(rb[q]);
    }
  };

  const int warp = tid >> 5, lane = tid & 31;
  if constexpr (TF32) {
    // 8 warps as 4 (i) x 2 (j): a warp owns 32 x 32 outputs = 2 x 4 mma tiles
    const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32, g = lane >> 2, t = lane & 3;
    float acc[2][4][4];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int z = 0; z < 4; ++z) acc[x][y][z] = 0.f;
    if (m_begin < m_end) load_stage(m_begin);
    for (int m0 = m_begin; m0 < m_end; m0 += kWgRowsStage) {
      __syncthreads();      // the previous stage's fragment loads are done
      store_stage();
      __syncthreads();
      if (m0 + kWgRowsStage < m_end) load_stage(m0 + kWgRowsStage);      // global loads of the next stage fly under this stage's MMAs
#pragma unroll
      for (int k8 = 0; k8 < kWgRowsStage; k8 += 8) {
        uint32_t af[2][4], bf[4][2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const float* base = sA + (k8 + t) * kWgLdA + wi + x * 16 + g;
          af[x][0] = __float_as_uint(base[0]);
          af[x][1] = __float_as_uint(base[8]);
          af[x][2] = __float_as_uint(base[4 * kWgLdA]);
          af[x][3] = __float_as_uint(base[4 * kWgLdA + 8]);
        }
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const float* base = sB + (k8 + t) * kWgLdB + wj + y * 8 + g;
          bf[y][0] = __float_as_uint(base[0]);
          bf[y][1] = __float_as_uint(base[4 * kWgLdB]);
        }
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) mma_tf32(acc[x][y], af[x][0], af[x][1], af[x][2], af[x][3], bf[y][0], bf[y][1]);
      }
    }
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          const int i = i0 + wi + x * 16 + g + (z >> 1) * 8, j = j0 + wj + y * 8 + 2 * t + (z & 1);
          if (i < p.na && j < p.nb && acc[x][y][z] != 0.f) atomicAdd(p.dw + (size_t)i * p.ld_dw + j, acc[x][y][z]);
        }
  } else {
    // exact fp32: thread = 8 (i) x 4 (j) outputs; 16 x 16 threads cover 128 x 64
    const int ti = (tid >> 4) * 8, tj = (tid & 15) * 4;
    float acc[8][4];
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
    if (m_begin < m_end) load_stage(m_begin);
    for (int m0 = m_begin; m0 < m_end; m0 += kWgRowsStage) {
      __syncthreads();
      store_stage();
      __syncthreads();
      if (m0 + kWgRowsStage < m_end) load_stage(m0 + kWgRowsStage);
#pragma unroll 4
      for (int k = 0; k < kWgRowsStage; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(sA + k * kWgLdA + ti), a1 = *reinterpret_cast<const float4*>(sA + k * kWgLdA + ti + 4);
        const float4 b = *reinterpret_cast<const float4*>(sB + k * kWgLdB + tj);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
      }
    }
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const int i = i0 + ti + x, j = j0 + tj + y;
        if (i < p.na && j < p.nb && acc[x][y] != 0.f) atomicAdd(p.dw + (size_t)i * p.ld_dw + j, acc[x][y]);
      }
  }
}

template <typename TA, typename TB>
static void wgrad_launch(const gvk_wgrad_params* p, dim3 grid, int rows_per_cta, cudaStream_t stream) {
  if (p->precision == GVK_PREC_TF32)
    wgrad_kernel<TA, TB, true><<<grid, kWgThreads, 0, stream>>>(*p, rows_per_cta);
  else
    wgrad_kernel<TA, TB, false><<<grid, kWgThreads, 0, stream>>>(*p, rows_per_cta);
}

int wgrad(const gvk_wgrad_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->a && p->b && p->dw, "gvk_wgrad: null pointer");
  GVK_CHECK_ARG(p->M > 0 && p->na > 0 && p->nb > 0, "gvk_wgrad: bad shape M=%d na=%d nb=%d", p->M, p->na, p->nb);
  GVK_CHECK_ARG(p->na % 4 == 0 && p->nb % 4 == 0 && p->lda % 4 == 0 && p->ldb % 4 == 0, "gvk_wgrad: na, nb, lda, ldb must be multiples of 4 (na=%d nb=%d lda=%d ldb=%d)",
                p->na, p->nb, p->lda, p->ldb);
  GVK_CHECK_ARG((p->a_dtype == GVK_F32 || p->a_dtype == GVK_BF16) && (p->b_dtype == GVK_F32 || p->b_dtype == GVK_BF16), "gvk_wgrad: dtypes must be GVK_F32 / GVK_BF16");
  GVK_CHECK_ARG((reinterpret_cast<uintptr_t>(p->a) & (p->a_dtype == GVK_F32 ? 15 : 7)) == 0 && (reinterpret_cast<uintptr_t>(p->b) & (p->b_dtype == GVK_F32 ? 15 : 7)) == 0,
                "gvk_wgrad: operands must be aligned to 4 elements");
  GVK_CHECK_ARG(p->precision == GVK_PREC_FP32 || p->precision == GVK_PREC_TF32, "gvk_wgrad: bad precision");
  GVK_CHECK_ARG(p->a_rows_per_batch >= 0 && p->b_rows_per_batch >= 0, "gvk_wgrad: negative rows_per_batch");
  const int ti = (p->na + kWgTA - 1) / kWgTA, tj = (p->nb + kWgTB - 1) / kWgTB;
  // rows split so that ~2 CTAs per SM exist; a CTA keeps at least 8 stages so the atomics stay a small share of its work
  int splits = std::max(1, (2 * sm_count() + ti * tj - 1) / (ti * tj));
  splits = std::min(splits, std::max(1, p->M / (8 * kWgRowsStage)));
  splits = std::min(splits, 65535);
  int rows_per_cta = (p->M + splits - 1) / splits;
  rows_per_cta = (rows_per_cta + kWgRowsStage - 1) / kWgRowsStage * kWgRowsStage;
  splits = (p->M + rows_per_cta - 1) / rows_per_cta;
  const dim3 grid(ti, tj, splits);
  if (p->a_dtype == GVK_F32 && p->b_dtype == GVK_F32) wgrad_launch<float, float>(p, grid, rows_per_cta, stream);
  else if (p->a_dtype == GVK_F32) wgrad_launch<float, __nv_bfloat16>(p, grid, rows_per_cta, stream);
  else if (p->b_dtype == GVK_F32) wgrad_launch<__nv_bfloat16, float>(p, grid, rows_per_cta, stream);
  else wgrad_launch<__nv_bfloat16, __nv_bfloat16>(p, grid, rows_per_cta, stream);
  GVK_CHECK_LAUNCH("wgrad");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// (2) H-frequency band removal + |.|
// ------------------------------------------------------------------------------------------------
constexpr int kHfRows = 40;       // output rows per CTA: 8 warps x 5 rows
constexpr int kHfCols = 160;      // output columns per CTA: lane owns columns 4 lane .. 4 lane + 3 and 128 + lane
constexpr int kHfThreads = 256;

__global__ void __launch_bounds__(kHfThreads, 1) hfreq_filter_kernel(gvk_hfreq_filter_params p) {
  extern __shared__ __align__(16) float hf_smem[];
  const int H = p.H, W = p.W;
  float* sX = hf_smem;                            // [H][kHfCols]
  float* sF = sX + (size_t)H * kHfCols;           // [H][8 warps][8]: F[k][row] for the CTA's 40 rows, 5 used of every 8
  const int slice = blockIdx.z, d = slice % p.D;
  const int r0 = blockIdx.y * kHfRows, c0 = blockIdx.x * kHfCols;
  const float* X = p.in + (size_t)slice * H * W;
  float* Y = p.out + (size_t)slice * H * W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!p.hit[d]) {      // pass-through slice: |x| on the CTA's tile
    for (int idx = tid; idx < kHfRows * kHfCols; idx += kHfThreads) {
      const int r = r0 + idx / kHfCols, c = c0 + idx % kHfCols;
      if (r < H && c < W) Y[(size_t)r * W + c] = fabsf(X[(size_t)r * W + c]);
    }
    return;
  }
  for (int idx = tid; idx < H * kHfCols; idx += kHfThreads) {
    const int k = idx / kHfCols, c = c0 + idx % kHfCols;
    sX[idx] = c < W ? X[(size_t)k * W + c] : 0.f;
  }
  for (int idx = tid; idx < H * 64; idx += kHfThreads) {
    const int k = idx >> 6, w = (idx >> 3) & 7, i = idx & 7, r = r0 + w * 5 + i;
    sF[idx] = (i < 5 && r < H) ? p.filt[(size_t)k * H + r] : 0.f;      // F is symmetric: row r of F read as column r
  }
  __syncthreads();
  float acc[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < H; ++k) {
    const float4 f4 = *reinterpret_cast<const float4*>(sF + (k * 8 + warp) * 8);
    const float f5 = sF[(k * 8 + warp) * 8 + 4];
    const float4 x4 = *reinterpret_cast<const float4*>(sX + k * kHfCols + lane * 4);
    const float x5 = sX[k * kHfCols + 128 + lane];
    const float fv[5] = {f4.x, f4.y, f4.z, f4.w, f5}, xv[5] = {x4.x, x4.y, x4.z, x4.w, x5};
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) acc[i][j] = fmaf(fv[i], xv[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int r = r0 + warp * 5 + i;
    if (r >= H) continue;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int c = c0 + (j < 4 ? lane * 4 + j : 128 + lane);
      if (c < W) Y[(size_t)r * W + c] = fabsf(acc[i][j]);
    }
  }
}

int hfreq_filter(const gvk_hfreq_filter_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->in && p->out && p->filt && p->hit, "gvk_hfreq_filter: null pointer");
  GVK_CHECK_ARG(p->slices > 0 && p->D > 0 && p->H > 0 && p->W > 0 && p->slices % p->D == 0, "gvk_hfreq_filter: bad shape slices=%d D=%d H=%d W=%d", p->slices, p->D, p->H, p->W);
  GVK_CHECK_ARG(p->in != p->out, "gvk_hfreq_filter: in-place is not supported (a CTA reads whole columns of the slice)");
  GVK_CHECK_ARG(p->slices <= 65535, "gvk_hfreq_filter: at most 65535 slices per call (got %d)", p->slices);
  const size_t smem = ((size_t)p->H * kHfCols + (size_t)p->H * 64) * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("gvk_hfreq_filter: H = %d needs %zu B of shared memory (> 227 KB)", p->H, smem);
    return GVK_ERR_UNSUPPORTED;
  }
  static size_t configured = 0;
  if (smem > configured) {
    int st = cuda_status(cudaFuncSetAttribute(hfreq_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "hfreq_filter smem");
    if (st != GVK_OK) return st;
    configured = smem;
  }
  const dim3 grid((p->W + kHfCols - 1) / kHfCols, (p->H + kHfRows - 1) / kHfRows, p->slices);
  hfreq_filter_kernel<<<grid, kHfThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("hfreq_filter");
  return GVK_OK;
}

}  // namespace gvk
