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
// A CTA owns a 128 x 64 tile of dw and a range of rows; 32-row stages of both operands travel through a 3-deep cp.async ring in their storage
// type (fp32 or bf16), so the loads of two stages are in flight under the MMAs of the third.  Fragments are converted when they are read: a
// bf16 value is a tf32 value as it stands (shift), an fp32 value is rounded to tf32 with one integer add (tf32_bits).
// ------------------------------------------------------------------------------------------------
constexpr int kWgTA = 128;     // tile extent along a's columns (rows of dw)
constexpr int kWgTB = 64;      // tile extent along b's columns (columns of dw)
constexpr int kWgRowsStage = 32;
constexpr int kWgStages = 3;
constexpr int kWgThreads = 256;
constexpr int kWgLdA = kWgTA + 8;   // row strides in elements: fragment loads (row t, column g) hit distinct banks (fp32) or pair up on words (bf16)
constexpr int kWgLdB = kWgTB + 8;

template <typename T>
__device__ __forceinline__ void cp_chunk4(T* smem_dst, const T* gsrc, bool valid);      // 4 elements, zero-filled when !valid
template <>
__device__ __forceinline__ void cp_chunk4<float>(float* smem_dst, const float* gsrc, bool valid) { cp_async16_zfill(smem_dst, gsrc, valid); }
template <>
__device__ __forceinline__ void cp_chunk4<__nv_bfloat16>(__nv_bfloat16* smem_dst, const __nv_bfloat16* gsrc, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(valid ? 8 : 0) : "memory");
}

template <bool TF32>
__device__ __forceinline__ uint32_t wg_frag(const float* p) { return TF32 ? tf32_bits(*p) : __float_as_uint(*p); }
template <bool TF32>
__device__ __forceinline__ uint32_t wg_frag(const __nv_bfloat16* p) { return (uint32_t)(*reinterpret_cast<const uint16_t*>(p)) << 16; }
__device__ __forceinline__ float wg_val(float v) { return v; }
__device__ __forceinline__ float wg_val(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ size_t wg_row(int m, int rpb, int batch_rows) {
  if (rpb <= 0) return (size_t)m;
  const int b = m / rpb;
  return (size_t)b * batch_rows + (m - b * rpb);
}

template <typename TA, typename TB>
constexpr size_t wg_smem_bytes() { return (size_t)kWgStages * kWgRowsStage * (kWgLdA * sizeof(TA) + kWgLdB * sizeof(TB)); }

template <typename TA, typename TB, bool TF32>
__global__ void __launch_bounds__(kWgThreads, 2) wgrad_kernel(gvk_wgrad_params p, int rows_per_cta) {
  extern __shared__ __align__(16) uint8_t wg_raw[];
  TA* sA = reinterpret_cast<TA*>(wg_raw);                                              // [stages][32][kWgLdA]
  TB* sB = reinterpret_cast<TB*>(wg_raw + (size_t)kWgStages * kWgRowsStage * kWgLdA * sizeof(TA));   // [stages][32][kWgLdB]
  const TA* A = reinterpret_cast<const TA*>(p.a);
  const TB* Bm = reinterpret_cast<const TB*>(p.b);
  const int i0 = blockIdx.x * kWgTA, j0 = blockIdx.y * kWgTB;
  const int m_begin = blockIdx.z * rows_per_cta, m_end = min(p.M, m_begin + rows_per_cta);
  const int nsteps = (m_end - m_begin + kWgRowsStage - 1) / kWgRowsStage;
  const int tid = threadIdx.x;
  // copy plan: A stage = 32 rows x 32 chunks of 4 elements -> 4 chunks per thread; B stage = 32 rows x 16 chunks -> 2 per thread
  auto issue = [&](int step) {
    if (step < nsteps) {
      const int m0 = m_begin + step * kWgRowsStage, slot = step % kWgStages;
      TA* dA = sA + (size_t)slot * kWgRowsStage * kWgLdA;
      TB* dB = sB + (size_t)slot * kWgRowsStage * kWgLdB;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = tid + q * kWgThreads, row = idx >> 5, c4 = (idx & 31) * 4;
        const int m = m0 + row, col = i0 + c4;
        const bool ok = m < m_end && col < p.na;
        cp_chunk4<TA>(dA + row * kWgLdA + c4, A + (ok ? wg_row(m, p.a_rows_per_batch, p.a_batch_rows) * p.lda + col : 0), ok);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int idx = tid + q * kWgThreads, row = idx >> 4, c4 = (idx & 15) * 4;
        const int m = m0 + row, col = j0 + c4;
        const bool ok = m < m_end && col < p.nb;
        cp_chunk4<TB>(dB + row * kWgLdB + c4, Bm + (ok ? wg_row(m, p.b_rows_per_batch, p.b_batch_rows) * p.ldb + col : 0), ok);
      }
    }
    cp_async_commit();
  };
  const int warp = tid >> 5, lane = tid & 31;
  // TF32: 8 warps as 4 (i) x 2 (j), a warp owns 32 x 32 outputs = 2 x 4 mma tiles.  FP32: thread = 8 (i) x 4 (j) outputs, 16 x 16 threads.
  const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32, g = lane >> 2, t = lane & 3;
  const int ti = (tid >> 4) * 8, tj = (tid & 15) * 4;
  float acc[32];
#pragma unroll
  for (int z = 0; z < 32; ++z) acc[z] = 0.f;
#pragma unroll
  for (int s0 = 0; s0 < kWgStages - 1; ++s0) issue(s0);
  for (int step = 0; step < nsteps; ++step) {
    cp_async_wait<kWgStages - 2>();
    __syncthreads();                       // stage `step` has landed for every thread; stage step - 1 has been read by every thread
    issue(step + kWgStages - 1);           // ... so its slot can be refilled
    const TA* cA = sA + (size_t)(step % kWgStages) * kWgRowsStage * kWgLdA;
    const TB* cB = sB + (size_t)(step % kWgStages) * kWgRowsStage * kWgLdB;
    if constexpr (TF32) {
#pragma unroll
      for (int k8 = 0; k8 < kWgRowsStage; k8 += 8) {
        uint32_t af[2][4], bf[4][2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const TA* base = cA + (k8 + t) * kWgLdA + wi + x * 16 + g;
          af[x][0] = wg_frag<true>(base);
          af[x][1] = wg_frag<true>(base + 8);
          af[x][2] = wg_frag<true>(base + 4 * kWgLdA);
          af[x][3] = wg_frag<true>(base + 4 * kWgLdA + 8);
        }
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const TB* base = cB + (k8 + t) * kWgLdB + wj + y * 8 + g;
          bf[y][0] = wg_frag<true>(base);
          bf[y][1] = wg_frag<true>(base + 4 * kWgLdB);
        }
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y)
            mma_tf32(*reinterpret_cast<float(*)[4]>(&acc[(x * 4 + y) * 4]), af[x][0], af[x][1], af[x][2], af[x][3], bf[y][0], bf[y][1]);
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < kWgRowsStage; ++k) {
        float av[8], bv[4];
#pragma unroll
        for (int x = 0; x < 8; ++x) av[x] = wg_val(cA[k * kWgLdA + ti + x]);
#pragma unroll
        for (int y = 0; y < 4; ++y) bv[y] = wg_val(cB[k * kWgLdB + tj + y]);
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x * 4 + y] = fmaf(av[x], bv[y], acc[x * 4 + y]);
      }
    }
  }
  cp_async_wait<0>();
  if constexpr (TF32) {
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          const int i = i0 + wi + x * 16 + g + (z >> 1) * 8, j = j0 + wj + y * 8 + 2 * t + (z & 1);
          const float v = acc[(x * 4 + y) * 4 + z];
          if (i < p.na && j < p.nb && v != 0.f) atomicAdd(p.dw + (size_t)i * p.ld_dw + j, v);
        }
  } else {
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const int i = i0 + ti + x, j = j0 + tj + y;
        if (i < p.na && j < p.nb && acc[x * 4 + y] != 0.f) atomicAdd(p.dw + (size_t)i * p.ld_dw + j, acc[x * 4 + y]);
      }
  }
}

template <typename TA, typename TB, bool TF32>
static int wgrad_launch2(const gvk_wgrad_params* p, dim3 grid, int rows_per_cta, cudaStream_t stream) {
  constexpr size_t smem = wg_smem_bytes<TA, TB>();
  static bool configured = false;
  if (!configured) {
    int st = cuda_status(cudaFuncSetAttribute(wgrad_kernel<TA, TB, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "wgrad smem");
    if (st != GVK_OK) return st;
    configured = true;
  }
  wgrad_kernel<TA, TB, TF32><<<grid, kWgThreads, smem, stream>>>(*p, rows_per_cta);
  return GVK_OK;
}

template <typename TA, typename TB>
static int wgrad_launch(const gvk_wgrad_params* p, dim3 grid, int rows_per_cta, cudaStream_t stream) {
  return p->precision == GVK_PREC_TF32 ? wgrad_launch2<TA, TB, true>(p, grid, rows_per_cta, stream) : wgrad_launch2<TA, TB, false>(p, grid, rows_per_cta, stream);
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
  int st;
  if (p->a_dtype == GVK_F32 && p->b_dtype == GVK_F32) st = wgrad_launch<float, float>(p, grid, rows_per_cta, stream);
  else if (p->a_dtype == GVK_F32) st = wgrad_launch<float, __nv_bfloat16>(p, grid, rows_per_cta, stream);
  else if (p->b_dtype == GVK_F32) st = wgrad_launch<__nv_bfloat16, float>(p, grid, rows_per_cta, stream);
  else st = wgrad_launch<__nv_bfloat16, __nv_bfloat16>(p, grid, rows_per_cta, stream);
  if (st != GVK_OK) return st;
  GVK_CHECK_LAUNCH("wgrad");
  return GVK_OK;
}

// ------------------------------------------------------------------------------------------------
// (2) H-frequency band removal + |.|
// ------------------------------------------------------------------------------------------------
constexpr int kHfRows = 40;       // output rows per pass: 8 warps x 5 rows
constexpr int kHfCols = 160;      // output columns per CTA: lane owns columns 4 lane .. 4 lane + 3 and 128 + lane
constexpr int kHfThreads = 256;

// A CTA owns one (slice, 160-column tile): the H x 160 tile of the slice is staged once (cp.async, every chunk in flight at once) and the CTA
// walks over the 40-row blocks of the output, the filter rows of the NEXT block loading under the FMAs of the current one.
__global__ void __launch_bounds__(kHfThreads, 1) hfreq_filter_kernel(gvk_hfreq_filter_params p) {
  extern __shared__ __align__(16) float hf_smem[];
  const int H = p.H, W = p.W;
  float* sX = hf_smem;                            // [H][kHfCols]
  float* sF = sX + (size_t)H * kHfCols;           // [2][H][kHfRows]: F[k][r0 + i] of the block being computed / being loaded (F is symmetric)
  const int slice = blockIdx.z, d = slice % p.D;
  const int c0 = blockIdx.x * kHfCols;
  const float* X = p.in + (size_t)slice * H * W;
  float* Y = p.out + (size_t)slice * H * W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.in) | reinterpret_cast<uintptr_t>(p.out)) & 15) == 0;
  if (!p.hit[d]) {      // pass-through slice: |x| on the CTA's columns
    if (vec) {
      const int w4 = min(kHfCols, W - c0) / 4;
      for (int idx = tid; idx < H * w4; idx += kHfThreads) {
        const int r = idx / w4, c = c0 + (idx - r * w4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(X + (size_t)r * W + c);
        *reinterpret_cast<float4*>(Y + (size_t)r * W + c) = make_float4(fabsf(v.x), fabsf(v.y), fabsf(v.z), fabsf(v.w));
      }
    } else {
      for (int idx = tid; idx < H * kHfCols; idx += kHfThreads) {
        const int r = idx / kHfCols, c = c0 + idx % kHfCols;
        if (c < W) Y[(size_t)r * W + c] = fabsf(X[(size_t)r * W + c]);
      }
    }
    return;
  }
  const bool fvec = (H % 4 == 0) && (reinterpret_cast<uintptr_t>(p.filt) & 15) == 0;
  auto stage_filter = [&](int rb, int buf) {      // sF[buf][k][i] = F[k][rb * 40 + i]  (rows past H: zero)
    float* dst = sF + (size_t)buf * H * kHfRows;
    const int r0 = rb * kHfRows;
    if (fvec) {
      for (int idx = tid; idx < H * (kHfRows / 4); idx += kHfThreads) {
        const int k = idx / (kHfRows / 4), i4 = (idx - k * (kHfRows / 4)) * 4;
        const bool ok = r0 + i4 < H;
        cp_async16_zfill(dst + k * kHfRows + i4, p.filt + (ok ? (size_t)k * H + r0 + i4 : 0), ok);
      }
    } else {
      for (int idx = tid; idx < H * kHfRows; idx += kHfThreads) {
        const int k = idx / kHfRows, i = idx - k * kHfRows;
        dst[idx] = r0 + i < H ? p.filt[(size_t)k * H + r0 + i] : 0.f;
      }
    }
    cp_async_commit();
  };
  if (vec) {
    for (int idx = tid; idx < H * (kHfCols / 4); idx += kHfThreads) {
      const int k = idx / (kHfCols / 4), c4 = (idx - k * (kHfCols / 4)) * 4;
      const bool ok = c0 + c4 < W;
      cp_async16_zfill(sX + k * kHfCols + c4, X + (ok ? (size_t)k * W + c0 + c4 : 0), ok);
    }
  } else {
    for (int idx = tid; idx < H * kHfCols; idx += kHfThreads) {
      const int k = idx / kHfCols, c = c0 + idx % kHfCols;
      sX[idx] = c < W ? X[(size_t)k * W + c] : 0.f;
    }
  }
  const int nrb = (H + kHfRows - 1) / kHfRows;
  stage_filter(0, 0);       // commits the slice tile's copies together with the first filter block
  for (int rb = 0; rb < nrb; ++rb) {
    if (rb + 1 < nrb) {
      stage_filter(rb + 1, (rb + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* cF = sF + (size_t)(rb & 1) * H * kHfRows + warp * 5;
    float acc[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float4 x4 = *reinterpret_cast<const float4*>(sX + k * kHfCols + lane * 4);
      const float x5 = sX[k * kHfCols + 128 + lane];
      const float xv[5] = {x4.x, x4.y, x4.z, x4.w, x5};
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float f = cF[k * kHfRows + i];      // same address for the whole warp: broadcast
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = fmaf(f, xv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int r = rb * kHfRows + warp * 5 + i;
      if (r >= H) continue;
      const int c = c0 + lane * 4;
      if (vec && c + 3 < W) {
        *reinterpret_cast<float4*>(Y + (size_t)r * W + c) = make_float4(fabsf(acc[i][0]), fabsf(acc[i][1]), fabsf(acc[i][2]), fabsf(acc[i][3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < W) Y[(size_t)r * W + c + j] = fabsf(acc[i][j]);
      }
      if (c0 + 128 + lane < W) Y[(size_t)r * W + c0 + 128 + lane] = fabsf(acc[i][4]);
    }
    __syncthreads();      // every warp is done with filter buffer rb & 1 before the copy of block rb + 2 overwrites it
  }
}

int hfreq_filter(const gvk_hfreq_filter_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->in && p->out && p->filt && p->hit, "gvk_hfreq_filter: null pointer");
  GVK_CHECK_ARG(p->slices > 0 && p->D > 0 && p->H > 0 && p->W > 0 && p->slices % p->D == 0, "gvk_hfreq_filter: bad shape slices=%d D=%d H=%d W=%d", p->slices, p->D, p->H, p->W);
  GVK_CHECK_ARG(p->in != p->out, "gvk_hfreq_filter: in-place is not supported (a CTA reads whole columns of the slice)");
  GVK_CHECK_ARG(p->slices <= 65535, "gvk_hfreq_filter: at most 65535 slices per call (got %d)", p->slices);
  const size_t smem = ((size_t)p->H * kHfCols + 2 * (size_t)p->H * kHfRows) * sizeof(float);
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
  const dim3 grid((p->W + kHfCols - 1) / kHfCols, 1, p->slices);
  hfreq_filter_kernel<<<grid, kHfThreads, smem, stream>>>(*p);
  GVK_CHECK_LAUNCH("hfreq_filter");
  return GVK_OK;
}

}  // namespace gvk
