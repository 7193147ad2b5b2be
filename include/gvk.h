/*
 * gvk.h — C ABI of libgvk_sm100a.so, the B200 (sm_100a) kernels behind gaviko_b200's drop-in GAViKO modules.
 *
 * Conventions (every entry point):
 *   - plain pointers + sizes, no torch types; every pointer is DEVICE memory owned by the caller unless it says "host";
 *   - no allocation, no ownership transfer, no hidden synchronisation: work is enqueued on `stream` and the call returns;
 *   - returns GVK_OK (0) or a negative gvk_status; gvk_last_error() gives the message (thread-local);
 *   - row-major matrices; `ld*` are leading dimensions in ELEMENTS;
 *   - dtype tags: GVK_F32 / GVK_BF16.
 *
 * "Replaces" citations are file:line of the reference (gMedAI-Lab/GAViKO, paths relative to src/).
 */
#ifndef GVK_H_
#define GVK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* gvk_stream_t; /* == cudaStream_t */

typedef enum {
  GVK_OK = 0,
  GVK_ERR_INVALID_ARGUMENT = -1,
  GVK_ERR_CUDA = -2,
  GVK_ERR_UNSUPPORTED = -3,
  GVK_ERR_NO_DEVICE = -4
} gvk_status;

enum { GVK_F32 = 0, GVK_BF16 = 1 };

/* activation selector of the GEMM epilogue */
enum { GVK_ACT_NONE = 0, GVK_ACT_GELU = 1, GVK_ACT_GELU_BWD = 2 };

const char* gvk_last_error(void);
int gvk_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t gvk_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Dense "TN" GEMM with fused epilogue:   acc[m,n] = sum_k A[m,k] * B[n,k]        (A: [M,K], B: [N,K], both K-contiguous)
 *     v = acc + bias[n]                                   (bias optional)
 *     v = v * ssf_scale[n] + ssf_shift[n]                 (optional; model/ssf.py:24-31)
 *     act == GELU     : if aux: aux[m,n] = v  (pre-activation, saved for backward);  v = gelu_erf(v)
 *     act == GELU_BWD : v = v * gelu_erf'(aux[m,n])
 *     v += pos[(m % rows_per_batch), n]                   (optional; positional embedding, model/gaviko.py:542-547)
 *     v += res1[m,n] + res2[m,n]                          (optional fp32 residuals)
 *     out [row(m), n] = v      row(m) = (m / rows_per_batch) * out_batch_rows + out_row_offset + m % rows_per_batch
 *     out2[m, n]      = v      (optional second copy, fp32; the GAViKO local token stream, model/gaviko.py:546-547)
 *
 * a_dtype == GVK_BF16: tcgen05/TMEM tensor-core kernel (TMA-fed, fp32 accumulate).   K % 64 == 0, lda/ldb % 8 == 0.
 * a_dtype == GVK_F32 : exact-fp32 FFMA kernel (the "fp32 mode" of the north star).
 *
 * Replaces: nn.Linear / F.linear calls of model/vision_transformer.py:31-35,53-58, the Conv3d patch embedding
 * model/gaviko.py:383-385,532-533 (as a GEMM over gathered patches) and their autograd dgrad counterparts.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a;
  const void* b;
  int ab_dtype; /* GVK_BF16 or GVK_F32 (both operands) */
  int M, N, K;
  int lda, ldb;

  const float* bias;
  const float* ssf_scale;
  const float* ssf_shift;
  int act;
  void* aux;
  int aux_dtype;
  int ld_aux;
  const float* pos;
  int rows_per_batch; /* 0 = no row remap / no pos */
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
} gvk_gemm_params;

int gvk_gemm(const gvk_gemm_params* p, gvk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GVK_H_ */
