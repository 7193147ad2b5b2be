// Shared device/host helpers for the gaviko_b200 sm_100a kernels.
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and small math helpers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvk.h"

namespace gvk {

// ------------------------------------------------------------------------------------------------
// host-side status plumbing
// ------------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);
int sm_count();

#define GVK_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      gvk::set_last_error(__VA_ARGS__);           \
      return GVK_ERR_INVALID_ARGUMENT;            \
    }                                             \
  } while (0)

void note_launch();
#define GVK_CHECK_LAUNCH(what)                                   \
  do {                                                           \
    gvk::note_launch();                                          \
    int _st = gvk::cuda_status(cudaGetLastError(), what);        \
    if (_st != GVK_OK) return _st;                               \
  } while (0)

// Tensor maps (driver entry point resolved lazily so the library loads without libcuda / a GPU).
int make_tma_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                     uint32_t box_rows, uint32_t box_cols);
int make_tma_3d_bf16(CUtensorMap* out, const void* base, uint64_t d2, uint64_t rows, uint64_t cols,
                     uint64_t ld_row_elems, uint64_t ld_d2_elems, uint32_t box_rows, uint32_t box_cols);

int make_tma_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);   // rank <= 5

// op dispatchers (one per extern "C" entry point; defined next to their kernels)
int gemm_dispatch(const gvk_gemm_params* p, cudaStream_t stream);
int layernorm_fwd(const gvk_layernorm_fwd_params* p, cudaStream_t stream);
int layernorm_bwd(const gvk_layernorm_bwd_params* p, cudaStream_t stream);
int rowproj_down(const gvk_rowproj_down_params* p, cudaStream_t stream);
int rowproj_up(const gvk_rowproj_up_params* p, cudaStream_t stream);
int skinny_wgrad(const gvk_skinny_wgrad_params* p, cudaStream_t stream);
size_t skinny_wgrad_ws_floats(int r, int dim, int M);
int rowproj_down_tc(const gvk_rowproj_down_params* p, cudaStream_t stream);
int rowproj_up_tc(const gvk_rowproj_up_params* p, cudaStream_t stream);
int skinny_wgrad_tc(const gvk_skinny_wgrad_params* p, cudaStream_t stream);
bool skinny_wgrad_tc_supported(const gvk_skinny_wgrad_params* p);
int layernorm_bwd_tc(const gvk_layernorm_bwd_params* p, cudaStream_t stream);
bool layernorm_bwd_tc_supported(const gvk_layernorm_bwd_params* p);
int layernorm_fwd_down(const gvk_layernorm_fwd_down_params* p, cudaStream_t stream);
int rowproj_up_down(const gvk_rowproj_up_down_params* p, cudaStream_t stream);
int small_wgrad(const float* a, int lda, int ra, const float* b, int ldb, int rb, int M, float* dw, cudaStream_t stream);
int small_matmul(const float* a, int lda, int ra, const float* w, int rb, int M, float* out, int ldo, cudaStream_t stream);
int colsum(const float* x, int ldx, int M, int dim, float* out, cudaStream_t stream);
int cast_f32_bf16(const float* x, int ldx, void* y, int ldy, int M, int dim, cudaStream_t stream);
int cast_bf16_f32(const void* x, int ldx, float* y, int ldy, int M, int dim, cudaStream_t stream);
int ssf_bwd(const gvk_ssf_bwd_params* p, cudaStream_t stream);
int dropout(const gvk_dropout_params* p, cudaStream_t stream);
int attn_simt_fwd(const gvk_attn_fwd_params* p, cudaStream_t stream);
int attn_simt_bwd(const gvk_attn_bwd_params* p, cudaStream_t stream);
bool attn_win_tc_supported(const gvk_attn_fwd_params* p);
int attn_win_tc_fwd(const gvk_attn_fwd_params* p, cudaStream_t stream);
int attn_win_tc_bwd(const gvk_attn_bwd_params* p, cudaStream_t stream);
int rescale_intensity(const gvk_rescale_intensity_params* p, cudaStream_t stream);
int split_pack_bf16(const float* src, int ld_src, int rows, int r, void* dst, int ld_dst, int width, int pattern, cudaStream_t stream);
int patch_gather(const float* img, int B, int C, int D, int H, int W, int fp, int ps, void* patches, int out_dtype, cudaStream_t stream);
int fill_rows(const float* a, const float* b, int R, int dim, float* out, int ld_out, int out_batch_rows, int out_row_offset, int B, cudaStream_t stream);
int batch_rowsum(const float* x, int ldx, int batch_rows, int row_offset, int R, int dim, int B, float* out, int accumulate, cudaStream_t stream);
int grad_sumsq(const float* grad, size_t n, float grad_scale, float* partials, cudaStream_t stream);
int clip_adam(float* param, const float* grad, float* m, float* v, size_t n, const float* partials, float max_norm, float grad_scale, float lr, float beta1,
              float beta2, float eps, float wd, int step, float* norm_out, cudaStream_t stream);
int patch_embed(const gvk_patch_embed_params* p, cudaStream_t stream);
int patch_embed_supported(const gvk_patch_embed_params* p);
int clip_adam_dyn(float* param, const float* grad, float* m, float* v, size_t n, const float* partials, float max_norm, float grad_scale, const float* lr_dev,
                  float beta1, float beta2, float eps, float wd, const long long* step_dev, float* norm_out, cudaStream_t stream);
int mhsa_fwd(const gvk_mhsa_fwd_params* p, cudaStream_t stream);
int mhsa_bwd(const gvk_mhsa_bwd_params* p, cudaStream_t stream);
int mhsa_ws_fwd(const gvk_mhsa_fwd_params* p, cudaStream_t stream);
int mhsa_fwd2(const gvk_mhsa_fwd_params* p, cudaStream_t stream);
int mhsa_bwd_ws(const gvk_mhsa_bwd_params* p, cudaStream_t stream);
int mhsa_bwd_pipe(const gvk_mhsa_bwd_params* p, cudaStream_t stream);
int debug_trace(uint32_t* out, int n_words);
uint32_t* trace_buffer();   // device buffer [kTraceRoles][kTraceN][2] (allocated on first use), or nullptr
size_t mhsa_bwd_ws_floats(int B, int T, int H);
size_t mhsa_bwd_mask_words(int B, int T, int H);
int prompt_fusion_fwd(const gvk_fusion_fwd_params* p, cudaStream_t stream);
int prompt_fusion_bwd(const gvk_fusion_bwd_params* p, cudaStream_t stream);
int quickgelu_bwd(const float* dy, const float* pre, float* y, size_t n, cudaStream_t stream);
int quickgelu_fwd(const float* x, float* y, size_t n, cudaStream_t stream);
int quickgelu_bwd_add(const float* dy, const float* pre, const float* res, float* y, size_t n, cudaStream_t stream);
int latent_xattn_fwd(const gvk_latent_xattn_fwd_params* p, cudaStream_t stream);
int latent_xattn_bwd(const gvk_latent_xattn_bwd_params* p, cudaStream_t stream);
int gate_scale(const float* x, const float* gate, float* y, size_t n, cudaStream_t stream);
int gate_grads(const float* x, const float* dy, const float* gate, float* dx, float* dgate, size_t n, cudaStream_t stream);
int relu_bwd(const float* dy, const float* z, float* y, size_t n, cudaStream_t stream);
int wgrad(const gvk_wgrad_params* p, cudaStream_t stream);
int hfreq_filter(const gvk_hfreq_filter_params* p, cudaStream_t stream);
int head_fwd(const gvk_head_fwd_params* p, cudaStream_t stream);
int head_bwd(const gvk_head_bwd_params* p, cudaStream_t stream);
int loss_fwd_bwd(const float* logits, const long long* target, int B, int C, int kind, float gamma, float eps, long long ignore_index, float* loss, float* dlogits,
                 cudaStream_t stream);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t elect.sync _|p, 0xffffffff;\n\t selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in tensor memory (K-major, 16-bit elements packed two per 32-bit column, lane = row).
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns, registers -> TMEM (thread i of the warp writes lane base_lane + i).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor (tcgen05), SWIZZLE_128B, "version 1" (Blackwell).
//   K-major operand tile  [rows][64 bf16] (128 B rows, 8-row swizzle atoms of 1024 B): SBO = 1024, LBO unused (=16 B).
//   MN-major operand tile [k rows][64 bf16 of the M/N extent]: SBO = 1024 (8 k-rows), LBO = byte stride between 64-wide MN panels.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32.  a_mn / b_mn = 1 selects an MN-major ("transposed") operand.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// Pull `bytes` starting at p (a whole number of 128-byte lines is touched) towards L2, spread over the 32 lanes of a warp.
// Fire-and-forget: costs no registers, lets a warp request a whole tile from HBM before it starts consuming it.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void warp_prefetch_l2(const void* p, int bytes, int lane) {
  const char* c = reinterpret_cast<const char*>(p);
  for (int off = lane * 128; off < bytes; off += 32 * 128) prefetch_l2(c + off);
}

// ---- math ----
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float fast_ex2(float x) {   // 2^x on the SFU, no denormal / range fix-up (softmax arguments are <= 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed fp32 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 work on an aligned register pair): one issue slot for two lanes of math.  The softmax
// warps of the attention kernels are bound by their own instruction stream, so halving the FMA / ADD count is worth more than the pipe rate.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float quick_gelu(float x) { return x * sigmoidf_(1.702f * x); }
__device__ __forceinline__ float quick_gelu_grad(float x) {
  const float s = sigmoidf_(1.702f * x);
  return s * (1.0f + 1.702f * x * (1.0f - s));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void st_from_float(T* p, float v);
template <>
__device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// Loads/stores with a runtime dtype tag (GVK_F32 / GVK_BF16).
__device__ __forceinline__ float ld_dyn(const void* base, size_t idx, int dtype) {
  return dtype == GVK_F32 ? reinterpret_cast<const float*>(base)[idx] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}
__device__ __forceinline__ void st_dyn(void* base, size_t idx, int dtype, float v) {
  if (dtype == GVK_F32)
    reinterpret_cast<float*>(base)[idx] = v;
  else
    reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
}

// cp.async (Ampere-style asynchronous global -> shared copies; used for small per-warp software pipelines)
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {   // !valid: 16 zero bytes, gsrc not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(valid ? 16 : 0) : "memory");
}

// tf32 mma.sync helpers (rank-r side paths and window attention in the bf16 compute mode)
__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float(f2tf32(x)); }
// Round-to-nearest (ties away) for an MMA operand in one integer add: the tensor core ignores the low 13 mantissa bits of a tf32 operand,
// so adding half an ulp of the 10-bit mantissa is all the rounding needs (finite inputs; the carry into the exponent is the right result).
__device__ __forceinline__ uint32_t tf32_bits(float x) { return __float_as_uint(x) + 0x1000u; }
// D(16x8) += A(16x8, row) * B(8x8, col).  lane = 4 g + t:  a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (k=t, n=g) b1 (k=t+4, n=g);
// d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1).
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Effective dropout seed: the by-value seed plus the device-resident replay counter (gvk.h, seed_salt).
__device__ __forceinline__ uint64_t salted_seed(uint64_t seed, const uint64_t* salt) { return salt ? seed + *salt * 0x9E3779B97F4A7C15ull : seed; }

// Philox4x32-10 (counter-based RNG for replayable dropout masks).
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// Attention-probability dropout of the tcgen05 MHSA kernels (gvk.h, gvk_mhsa_fwd_params): one Philox call decides 16 consecutive keys of a
// query row with 8-bit thresholds.  Returns bit e = 1 iff key 16 * k16 + e of query row q is KEPT.
struct MhsaDrop {
  uint32_t thr4;     // the 8-bit keep threshold replicated into the four bytes
  uint2 key;
  float inv_keep;    // 256 / threshold
  const uint64_t* salt;   // optional device-resident replay counter (gvk.h, seed_salt)
  unsigned long long seed;
};
__device__ __forceinline__ MhsaDrop mhsa_salted(MhsaDrop d) {      // once per thread, before the first mhsa_keep16
  const uint64_t s = salted_seed(d.seed, d.salt);
  d.key = make_uint2(static_cast<uint32_t>(s), static_cast<uint32_t>(s >> 32));
  return d;
}
__host__ inline MhsaDrop make_mhsa_drop(float drop_p, unsigned long long seed, const uint64_t* salt = nullptr) {
  MhsaDrop d;
  d.salt = salt;
  d.seed = seed;
  int thr = static_cast<int>(256.0f * (1.0f - drop_p) + 0.5f);
  thr = thr < 1 ? 1 : (thr > 256 ? 256 : thr);
  d.thr4 = thr >= 256 ? 0u : static_cast<uint32_t>(thr) * 0x01010101u;   // 0: keep everything (drop_p rounds to 0)
  d.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  d.inv_keep = 256.0f / static_cast<float>(thr);
  return d;
}
__device__ __forceinline__ uint32_t mhsa_keep16(const MhsaDrop& d, uint32_t bh, uint32_t q, uint32_t k16) {
  if (d.thr4 == 0u) return 0xFFFFu;
  const uint4 r = philox4x32(make_uint4(q, k16, bh, 0x6d687361u), d.key);
  // per byte: r < thr  <=>  the carry out of r + (256 - thr) is clear; __vcmpltu4 gives 0xFF per true byte, the multiply gathers the four flags
  auto keep4 = [&](uint32_t w) { return ((__vcmpltu4(w, d.thr4) & 0x01010101u) * 0x01020408u) >> 24; };
  return (keep4(r.x) & 15u) | ((keep4(r.y) & 15u) << 4) | ((keep4(r.z) & 15u) << 8) | ((keep4(r.w) & 15u) << 12);
}
// Timeline trace for tuning the attention kernels: CTA 0 records (tag, clock) pairs per role; read back with gvk_debug_trace (tools/mhsa_trace.py).
constexpr int kTraceN = 2048, kTraceRoles = 4;
struct Tracer {
  uint2* base;
  int n;
  __device__ __forceinline__ void init(uint32_t* buf, int role, bool on) { base = (on && buf && blockIdx.x == 0) ? reinterpret_cast<uint2*>(buf) + role * kTraceN : nullptr; n = 0; }
  __device__ __forceinline__ void operator()(uint32_t tag) {
    if (base && n < kTraceN) base[n++] = make_uint2(tag, static_cast<uint32_t>(clock()));
  }
};
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)
#endif  // __CUDACC__

}  // namespace gvk
