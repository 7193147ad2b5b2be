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
enum { GVK_ACT_NONE = 0, GVK_ACT_GELU = 1, GVK_ACT_GELU_BWD = 2, GVK_ACT_GELU_SAVE_GRAD = 3, GVK_ACT_MUL_AUX = 4 };

const char* gvk_last_error(void);
int gvk_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t gvk_launch_count(void);
/* sizeof() of a parameter struct by its typedef name (-1 if unknown): lets a binding verify its mirror of this header. */
long long gvk_struct_size(const char* name);

/* ------------------------------------------------------------------------------------------------------------------
 * Dense "TN" GEMM with fused epilogue:   acc[m,n] = sum_k A[m,k] * B[n,k]        (A: [M,K], B: [N,K], both K-contiguous)
 *     v = acc + bias[n]                                   (bias optional)
 *     v = v * ssf_scale[n] + ssf_shift[n]                 (optional; model/ssf.py:24-31)
 *     act == GELU     : if aux: aux[m,n] = v  (pre-activation, saved for backward);  v = gelu_erf(v)
 *     act == GELU_BWD : v = v * gelu_erf'(aux[m,n])
 *     act == GELU_SAVE_GRAD : aux[m,n] = gelu_erf'(v) (the derivative, which shares its erf / exp with the activation);  v = gelu_erf(v)
 *     act == MUL_AUX  : v = v * aux[m,n]           (backward of GELU_SAVE_GRAD: the dgrad epilogue is one multiply instead of an erf)
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


/* ------------------------------------------------------------------------------------------------------------------
 * Row kernels (a warp owns 1-4 token rows at a time; fp32 math; HBM-bound).  dim in {192, 384, 768, 1024}; rank r <= 96 for the
 * projections (as far as the [r, dim] fp32 panel fits 227 KB of shared memory), r <= 32 for weight gradients and LayerNorm backward.
 * Strided weights: element (j, c) of a rank-r projection is w[j * w_sj + c * w_sc], so an nn.Linear(dim, r).weight
 * ([r, dim]) is (w_sj = dim, w_sc = 1) and an nn.Linear(r, dim).weight ([dim, r]) used transposed is (w_sj = 1, w_sc = r).
 * seed_salt (optional, every struct with a seed): device pointer to a 64-bit counter mixed into the seed INSIDE the kernel
 * (seed + *seed_salt * 0x9E3779B97F4A7C15), so that a CUDA graph that baked `seed` into its kernel nodes still draws fresh masks on every replay
 * (the graph increments the counter, gaviko_b200/graph.py); forward and backward of one step must see the same counter value.
 * Dropout (replayable): element (m, c) of an [M, dim] tensor is kept iff philox(seed, offset + m * dim + c) >= drop_p and
 * scaled by 1 / (1 - drop_p).  `offset` must be a multiple of 4.
 * ------------------------------------------------------------------------------------------------------------------ */

/* y = LayerNorm(x) * gamma + beta  (eps inside the sqrt, biased variance — nn.LayerNorm; model/vision_transformer.py:30,49).
 * Optional SSF scale/shift after the norm (model/ssf.py:65,102).  mean / rstd (optional, [M]) are saved for backward. */
typedef struct {
  const float* x; int ldx;
  const float* gamma; const float* beta; float eps;
  const float* ssf_scale; const float* ssf_shift;
  void* y; int y_dtype; int ldy;
  float* mean; float* rstd;
  int M, dim;
} gvk_layernorm_fwd_params;
int gvk_layernorm_fwd(const gvk_layernorm_fwd_params* p, gvk_stream_t stream);

enum { GVK_ROWACT_NONE = 0, GVK_ROWACT_QUICKGELU = 1, GVK_ROWACT_RELU = 2 };

/* Arithmetic of the rank-r products in gvk_rowproj_down / gvk_rowproj_up / gvk_skinny_wgrad (`precision` field):
 *   GVK_PREC_FP32: exact fp32 FMAs (the 1e-4 parity mode);
 *   GVK_PREC_TF32: operands rounded to tf32 (10-bit mantissa), fp32 accumulation, on the tensor cores — used by the bf16 compute mode,
 *                  where these kernels then run at the HBM roofline instead of the FMA-issue limit. */
enum { GVK_PREC_FP32 = 0, GVK_PREC_TF32 = 1 };

/* z[m, j] = act( sum_c f(x[m, c]) * w(j, c) + bias[j] ),  f = optional dropout mask then optional LayerNorm.
 * pre (optional) receives the pre-activation.  Optional chained projection z2[m, k] = sum_j z[m, j] * w2[k * r + j]  (r2 <= 96).
 * Replaces LocalSelfAttention.norm/proj_down/qkv (model/gaviko.py:231-232), Awakening_Prompt.proj_down (model/gaviko.py:155-156)
 * and, with transposed strides, the dgrad of every rank-r up-projection. */
typedef struct {
  const float* x; int ldx; int M, dim, r;
  const float* ln_gamma; const float* ln_beta; float eps; float* mean; float* rstd;
  const float* w; int w_sj, w_sc; const float* bias; int act;
  float* pre; float* z; int ldz;
  const float* w2; int r2; float* z2; int ldz2;
  float drop_p; uint64_t seed; uint64_t offset; const uint64_t* seed_salt;
  int precision;
} gvk_rowproj_down_params;
int gvk_rowproj_down(const gvk_rowproj_down_params* p, gvk_stream_t stream);

/* LayerNorm forward and a rank-r down-projection of the RAW rows in one pass over x (bf16 compute mode):
 *   y = LayerNorm(x) * gamma + beta  (bf16; mean / rstd optional),   z = act(x W^T + bias),  pre (optional) = the pre-activation.
 * The two readers of the residual stream after attention in a GAViKO layer: FeedForward's norm (model/vision_transformer.py:30, called at
 * model/gaviko.py:304) and Awakening_Prompt.proj_down (model/gaviko.py:155-156).  The rank-r product runs as tf32 mma.sync (GVK_PREC_TF32
 * arithmetic); dim 384 or 768, r <= 32, w(j, c) strided like gvk_rowproj_down's. */
typedef struct {
  const float* x; int ldx; int M, dim;
  const float* gamma; const float* beta; float eps;
  void* y; int ldy;
  float* mean; float* rstd;
  const float* w; int w_sj, w_sc; const float* bias; int r; int act;
  float* pre; float* z; int ldz;
} gvk_layernorm_fwd_down_params;
int gvk_layernorm_fwd_down(const gvk_layernorm_fwd_down_params* p, gvk_stream_t stream);

/* out[m, c] = res[m, c] + dropout( sum_j c[m, j] * w(j, c) + bias[c] );  out_lp is an optional bf16 copy of out.
 * Replaces LocalSelfAttention.proj_up + proj_drop + residual (model/gaviko.py:242-243, 301), Awakening_Prompt.proj_up
 * (model/gaviko.py:187) and, with transposed strides, the dgrad of every rank-r down-projection. */
typedef struct {
  const float* c; int ldc; int M, dim, r;
  const float* w; int w_sj, w_sc; const float* bias;
  const float* res; int ld_res;
  float* out; int ld_out; void* out_lp; int ld_out_lp;
  float drop_p; uint64_t seed; uint64_t offset; const uint64_t* seed_salt;
  int precision;
} gvk_rowproj_up_params;
int gvk_rowproj_up(const gvk_rowproj_up_params* p, gvk_stream_t stream);

/* gvk_rowproj_up followed by gvk_rowproj_down on the rows it has just produced, in one pass over the [M, dim] stream (bf16 compute mode,
 * tf32 tensor-core arithmetic for both rank-r products; dim 384 or 768, r and r2 <= 24):
 *   out[m, c] = res[m, c] + drop_up( sum_j c[m, j] * w(j, c) + bias[c] )
 *   z[m, k]   = act( sum_c drop_dn(out[m, c]) * w2(k, c) + bias2[k] ),   pre (optional) = the pre-activation.
 * Forward of a GAViKO layer: LocalSelfAttention.proj_up + proj_drop + residual (model/gaviko.py:242-243, 301), then Awakening_Prompt.proj_down
 * of the new local stream (model/gaviko.py:155-156).  Backward: d(loc) += d(ul) Wd, then the dgrad of proj_up with the replayed proj_drop
 * mask.  Both dropout masks follow gvk_rowproj_up's rule (element index = offset + m * dim + c; offsets multiples of 4).  out may alias res. */
typedef struct {
  const float* c; int ldc; int M, dim, r;
  const float* w; int w_sj, w_sc; const float* bias;
  const float* res; int ld_res;
  float* out; int ld_out;
  float up_drop_p; uint64_t up_seed; uint64_t up_offset;
  const float* w2; int w2_sj, w2_sc; const float* bias2; int r2; int act;
  float* pre; float* z; int ldz;
  float dn_drop_p; uint64_t dn_seed; uint64_t dn_offset;
  const uint64_t* seed_salt;
} gvk_rowproj_up_down_params;
int gvk_rowproj_up_down(const gvk_rowproj_up_down_params* p, gvk_stream_t stream);

/* Rank-r weight gradient:  dw(j, c) += sum_m a[m, j] * f(x[m, c]);  da_colsum[j] += sum_m a[m, j];  dx_colsum[c] += sum_m f(x[m, c]).
 * f = optional dropout mask, then optional LayerNorm recomputed from saved mean / rstd.  Outputs ACCUMULATE (zero them first).
 * Deterministic two-stage reduction (per-CTA partials in `ws`, then one reduce launch): `ws` must hold at least
 * gvk_skinny_wgrad_ws_floats(r, dim, M) floats. */
size_t gvk_skinny_wgrad_ws_floats(int r, int dim, int M);
typedef struct {
  const float* a; int lda; int r;
  const float* x; int ldx; int dim; int M;
  const float* ln_gamma; const float* ln_beta; const float* mean; const float* rstd;
  float* dw; int dw_sj, dw_sc;
  float* da_colsum; float* dx_colsum;
  float drop_p; uint64_t seed; uint64_t offset; const uint64_t* seed_salt;
  float* ws; size_t ws_floats;
  int precision;
} gvk_skinny_wgrad_params;
int gvk_skinny_wgrad(const gvk_skinny_wgrad_params* p, gvk_stream_t stream);

/* LayerNorm backward:  dx = dres + LN'(dy) + az @ aw;  dy = dense ([M, dim] fp32) and / or rank-r (sum_j dz[m, j] * w(j, c)), added together.
 * With ssf_scale the forward was y = LN(x) * ssf_scale + ssf_shift (model/ssf.py:65,102): the incoming gradient is first reduced into
 * dssf_scale += colsum(dy * LN(x)), dssf_shift += colsum(dy) (needs beta) and multiplied by ssf_scale.
 * az @ aw (optional, az [M, ra], aw(j, c) strided like w) is added OUTSIDE the norm: the dgrad of a rank-ra down-projection that reads
 * the same residual stream as the LayerNorm (Awakening_Prompt.proj_down next to FeedForward's norm, model/gaviko.py:155,304).
 * dgamma / dbeta (optional, [dim]) accumulate with atomics.  dx may alias dres or dy.
 * dx_lp is an optional bf16 copy (the next dgrad GEMM's A operand).  dx may alias dy only when dy is fp32.
 * precision = GVK_PREC_TF32 (bf16 compute mode) runs the rank-r product (az @ aw next to a dense bf16 dy, or dz @ w with dgamma / dbeta)
 * as mma.sync m16n8k8 with tf32 operands inside the same pass; the LayerNorm arithmetic itself stays fp32.  Forms outside those two run
 * the exact kernel whatever the field says. */
typedef struct {
  const void* dy; int ld_dy; int dy_dtype;   /* GVK_F32 or GVK_BF16 (the dgrad GEMM that produces dy then writes half the bytes) */
  const float* dz; int ld_dz; const float* w; int w_sj, w_sc; int r;
  const float* x; int ldx; const float* gamma; const float* mean; const float* rstd;
  const float* dres; int ld_dres;
  float* dx; int ld_dx; void* dx_lp; int ld_dx_lp;
  float* dgamma; float* dbeta;
  int M, dim;
  const float* az; int ld_az; const float* aw; int aw_sj, aw_sc; int ra;
  const float* beta; const float* ssf_scale; float* dssf_scale; float* dssf_shift;
  int precision;                /* GVK_PREC_FP32 (exact, default) or GVK_PREC_TF32 */
  /* optional projection of the OUTPUT rows, oz[m, j] = sum_c dx[m, c] * ow(j, c)  (ow strided like w; orank <= 24): the dgrad of the rank-r
   * up-projection that reads the residual gradient next (Awakening_Prompt.proj_up, model/gaviko.py:187).  GVK_PREC_TF32 with a dense bf16
   * dy and no other rank term only; any other request with ow set is refused. */
  const float* ow; int ow_sj, ow_sc; int orank; float* oz; int ld_oz;
} gvk_layernorm_bwd_params;
int gvk_layernorm_bwd(const gvk_layernorm_bwd_params* p, gvk_stream_t stream);

/* dw[j * rb + k] += sum_m a[m, j] * b[m, k]   (ra, rb <= 96, ra*rb <= 4096; the LocalSelfAttention.qkv weight gradient). */
int gvk_small_wgrad(const float* a, int lda, int ra, const float* b, int ldb, int rb, int M, float* dw, gvk_stream_t stream);

/* out[m, k] = sum_j a[m, j] * w[j * rb + k]   (ra, rb <= 96, ra*rb <= 4096; dgrad of the LocalSelfAttention.qkv projection). */
int gvk_small_matmul(const float* a, int lda, int ra, const float* w, int rb, int M, float* out, int ldo, gvk_stream_t stream);

/* y[m, n] += / = colsum helpers: out[c] += sum_m x[m, c]  (bias gradients; bitfit). */
int gvk_colsum(const float* x, int ldx, int M, int dim, float* out, gvk_stream_t stream);

/* Elementwise cast fp32 -> bf16 of an [M, dim] matrix (weights, activations), and back. */
int gvk_cast_f32_bf16(const float* x, int ldx, void* y, int ldy, int M, int dim, gvk_stream_t stream);
int gvk_cast_bf16_f32(const void* x, int ldx, float* y, int ldy, int M, int dim, gvk_stream_t stream);

/* Backward of an SSF site y = x * scale + shift (model/ssf.py:24-31) and of plain bias adds (bitfit), over logical rows m < M whose
 * physical row is (m / rows_per_batch) * batch_rows + m % rows_per_batch in dy, y and dx (rows_per_batch == 0: identity):
 *   dshift[n] += sum_m dy[m, n];   dscale[n] += sum_m dy[m, n] * x[m, n];   dx[m, n] = dy[m, n] * scale[n]
 * The site's INPUT x is recovered from its saved OUTPUT: x = (y - sub[m % rows_per_batch, n] - shift[n]) / scale[n]  (`sub` = optional
 * term added after the site, e.g. the positional embedding after the patch-embedding site, model/ssf.py:236-240).
 * scale == NULL: bias-gradient mode (only dshift; y / dx unused).  dy, y, dx share `dtype` (GVK_F32 / GVK_BF16); dx may alias dy. */
typedef struct {
  const void* dy; int ld_dy; const void* y; int ld_y; void* dx; int ld_dx; int dtype;
  const float* scale; const float* shift; const float* sub; int ld_sub;
  float* dscale; float* dshift;
  int M, N, rows_per_batch, batch_rows;
} gvk_ssf_bwd_params;
int gvk_ssf_bwd(const gvk_ssf_bwd_params* p, gvk_stream_t stream);

/* out = res + dropout(x) elementwise over an [M, N] matrix (x / out in `dtype`, res optional fp32 [M, N] with out then fp32... see below).
 * Replayable mask: element e = offset + m * N + n is kept iff byte (e % 16) of philox4x32-10(counter = (e / 16 low, e / 16 high, 0, 'drop'),
 * key = seed [+ seed_salt]) is < thr = round(256 (1 - drop_p)); kept values are scaled by 256 / thr (unbiased for the drop probability
 * actually applied, 1 - thr / 256 — the rule of the attention-probability dropout).  One Philox call decides 16 elements, which keeps the
 * pass HBM-bound.  N % 4 == 0, offset % 4 == 0.  (The rank-r kernels keep their own rule: one 32-bit word per element.)
 * x_dtype / out_dtype in {GVK_F32, GVK_BF16}; res (optional) is fp32.  Used for the nn.Dropout sites of the un-frozen train mode
 * (model/vision_transformer.py:33,35,58 and :157) and, applied to gradients with the same seed, for their backward. */
typedef struct {
  const void* x; int x_dtype; int ldx; const float* res; int ld_res; void* out; int out_dtype; int ld_out;
  int M, N; float drop_p; uint64_t seed; uint64_t offset; const uint64_t* seed_salt;
} gvk_dropout_params;
int gvk_dropout(const gvk_dropout_params* p, gvk_stream_t stream);


/* ------------------------------------------------------------------------------------------------------------------
 * Softmax attention, CUDA-core (SIMT) implementation: exact fp32 math, one warp per query (forward, dQ) or per key (dK, dV),
 * no atomics, nothing of size T x T is ever materialised.  Used for
 *   (1) GAViKO's window-sparse LocalSelfAttention core (model/gaviko.py:235-241): D = local_dim (20), H = 1, additive
 *       {0,-inf} window mask given by its closed form (allowed j: i_ax - k_ax/2 <= j_ax <= i_ax + k_ax - 1 - k_ax/2 per axis,
 *       model/gaviko.py:212-227), dropout on the probabilities;
 *   (2) the frozen MHSA core in fp32 mode (model/vision_transformer.py:65-70): D = 64, dense.
 * Layout: rows are tokens (B*T rows); head h of q / k / v lives at columns q_off / k_off / v_off + h*D of `qkv`.
 * Dropout element index of probability (b, h, i, j) is ((b*H + h)*T + i)*T + j  (philox(seed, offset + index)).
 * precision = GVK_PREC_TF32 (bf16 compute mode) runs case (1) on the tensor cores instead: mma.sync m16n8k8 with tf32 operands,
 * fp32 accumulation and softmax, K / V rows of the chunk's window staged once in shared memory (fp32 rows, D in {20, 32},
 * grid_d + grid_h + grid_w <= 31, the staged window must fit in shared memory; other problems silently use the exact path).  Its dropout
 * mask is drawn per aligned 2x2 (query, key) block, so forward and backward must use the same precision.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* qkv; int dtype; int ld;
  int q_off, k_off, v_off;
  int B, T, H, D;
  float scale;
  int win_d, win_h, win_w;      /* window extents; win_d == 0 selects dense attention */
  int grid_d, grid_h, grid_w;   /* token grid (T == grid_d*grid_h*grid_w when windowed) */
  float drop_p; uint64_t seed; uint64_t offset; const uint64_t* seed_salt;
  void* out; int ld_out;        /* [B*T, H*D], same dtype as qkv */
  float* lse;                   /* [B*H*T] log-sum-exp of the scaled scores (saved for backward) */
  int precision;                /* GVK_PREC_FP32 (exact, default) or GVK_PREC_TF32 (windowed case only) */
} gvk_attn_fwd_params;
int gvk_attn_simt_fwd(const gvk_attn_fwd_params* p, gvk_stream_t stream);

typedef struct {
  gvk_attn_fwd_params f;        /* same problem description as forward (out / lse are inputs here) */
  const void* dout; int ld_dout; /* [B*T, H*D] */
  float* delta;                 /* workspace [B*H*T] */
  void* dqkv; int ld_dqkv;      /* gradient in the layout of qkv (q/k/v column blocks are fully overwritten) */
} gvk_attn_bwd_params;
int gvk_attn_simt_bwd(const gvk_attn_bwd_params* p, gvk_stream_t stream);


/* ------------------------------------------------------------------------------------------------------------------
 * Frozen multi-head self-attention core on the tensor cores (bf16 operands, fp32 accumulation / softmax), flash style:
 * TMA-staged Q/K/V tiles, S = Q K^T and O += P V as tcgen05.mma with TMEM accumulators, online softmax in registers,
 * nothing of size T x T touches HBM.  Handles the prompt-extended ragged sequence (any T; tails are masked in-kernel).
 * Layout: qkv [B*T, 3*H*64] bf16, columns [q | k | v], head-major inside each block (einops 'b n (h d) -> b h n d',
 * model/vision_transformer.py:62-63); out [B*T, H*64] bf16; lse [B*H*T] fp32 (natural-log, of the scaled scores).
 * Replaces model/vision_transformer.py:65-71 and its autograd backward.  head dim is fixed at 64.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* qkv; int ld;
  int B, T, H;
  float scale;
  void* out; int ld_out;
  float* lse;
  /* Attention-probability dropout (model/vision_transformer.py:50,69; active for --method melo / linear / bitfit in train mode): probability
   * (b, h, i, j) is kept iff byte (j % 16) of philox4x32-10(counter = (i, j / 16, b*H + h, 'mhsa'), key = seed) < round(256 (1 - drop_p)); the
   * kept ones are scaled by 256 / round(256 (1 - drop_p)).  lse stays that of the un-dropped softmax.  drop_p = 0: no dropout. */
  float drop_p; uint64_t seed; const uint64_t* seed_salt;
} gvk_mhsa_fwd_params;
int gvk_mhsa_fwd(const gvk_mhsa_fwd_params* p, gvk_stream_t stream);

typedef struct {
  const void* qkv; int ld;
  int B, T, H;
  float scale;
  const void* out; int ld_out;
  const float* lse;
  const void* dout; int ld_dout;   /* [B*T, H*64] bf16 */
  float* delta;                    /* workspace of gvk_mhsa_bwd_ws_floats(B, T, H) floats (16-byte aligned): rowsum(dO*O) and the log2-domain lse handed
                                      from the dQ to the dK/dV kernel, rows padded to the 128-row blocks of the kernels */
  void* dqkv; int ld_dqkv;         /* [B*T, 3*H*64] bf16, fully overwritten */
  float drop_p; uint64_t seed; const uint64_t* seed_salt;     /* the forward call's dropout: the mask is regenerated (dQ kernel) and handed to the dK/dV kernel through mask_ws */
  uint32_t* mask_ws;               /* drop_p > 0: workspace of gvk_mhsa_bwd_mask_words(B, T, H) 32-bit words (16-byte aligned); else unused */
} gvk_mhsa_bwd_params;
int gvk_mhsa_bwd(const gvk_mhsa_bwd_params* p, gvk_stream_t stream);
size_t gvk_mhsa_bwd_ws_floats(int B, int T, int H);
size_t gvk_mhsa_bwd_mask_words(int B, int T, int H);
/* Tuning aid: with GVK_PIPE_DBG & 4 the backward kernels record a (tag, clock) timeline of CTA 0; copies up to n_words 32-bit words of it to the host. */
int gvk_debug_trace(uint32_t* out, int n_words);

/* ------------------------------------------------------------------------------------------------------------------
 * Token assembly (a1/a2 of the hot path)
 * ------------------------------------------------------------------------------------------------------------------ */

/* Per-volume intensity rescale in front of the path: torchio.RescaleIntensity(out_min_max=(lo, hi)) with its default
 * percentiles (0, 100), the last transform of every pipeline of the reference (train.py:53,57,61; eval.py:31; inference.py:30), there run on
 * the CPU in the DataLoader workers.  For each of the B volumes of n contiguous fp32 values:
 *   y = (x - min) / (max - min) * (hi - lo) + lo      (fp32, IEEE division, this order of operations: bit-identical to the CPU transform);
 * a constant volume is passed through unchanged (torchio warns and returns its input).  out may alias in when out_dtype is GVK_F32.
 * ws: 2 * GVK_RESCALE_PARTS * B floats of workspace (per-volume min / max partials; no atomics, deterministic). */
#define GVK_RESCALE_PARTS 64
typedef struct {
  const float* in; void* out; int out_dtype;
  int B; long long n;
  float out_min, out_max;
  float* ws;
} gvk_rescale_intensity_params;
int gvk_rescale_intensity(const gvk_rescale_intensity_params* p, gvk_stream_t stream);

/* Non-overlapping 3-D patch gather in Conv3d weight order (model/gaviko.py:383-385,532-533):
 *   patches[b*N + (d*nh + h)*nw + w, ((c*fp + kd)*ps + kh)*ps + kw] = img[b, c, d*fp + kd, h*ps + kh, w*ps + kw]
 * img fp32 contiguous (B, C, D, H, W); patches row-major [B*N, C*fp*ps*ps] in out_dtype. */
int gvk_patch_gather(const float* img, int B, int C, int D, int H, int W, int fp, int ps, void* patches, int out_dtype, gvk_stream_t stream);

/* Patch embedding in one kernel (a1 + a2 of the hot path; replaces Gaviko.conv_proj = nn.Conv3d(C, dim, kernel = stride = (fp, ps, ps)),
 * flatten(2).transpose(1, 2) and the patch rows of the token assembly, model/gaviko.py:383-385,532-548): 5-D TMA boxes gather the
 * non-overlapping patches of the fp32 volume straight into the operand ring of a tcgen05 GEMM (tf32 operands, fp32 accumulation), whose
 * epilogue writes   x = conv(img)[b, tok, :] + bias + pos[tok, :]   to  out[b*out_batch_rows + out_row_offset + tok, :]  and, when out2 is
 * set, to out2[b*n_tok + tok, :].  tok = (d*gh + h)*gw + w.  weight is the Conv3d weight as stored: [dim, C*fp*ps*ps] fp32.
 * gvk_patch_embed_supported() != 0 for ps = 16 and <= 256 tokens per (b, d) plane; other geometries use gvk_patch_gather + gvk_gemm. */
typedef struct {
  const float* img; int B, C, D, H, W; int fp, ps;
  const float* weight;
  const float* bias;
  const float* pos;
  int dim;
  float* out; int ld_out; int out_batch_rows, out_row_offset;
  float* out2; int ld_out2;
} gvk_patch_embed_params;
int gvk_patch_embed(const gvk_patch_embed_params* p, gvk_stream_t stream);
int gvk_patch_embed_supported(const gvk_patch_embed_params* p);

/* out[b*out_batch_rows + out_row_offset + r, :] = a[r, :] + b[r, :]   for b < B, r < R   (b may be NULL).
 * Writes the batch-broadcast prompt / cls rows of the token matrix (model/gaviko.py:536-543). */
int gvk_fill_rows(const float* a, const float* b, int R, int dim, float* out, int ld_out, int out_batch_rows, int out_row_offset, int B, gvk_stream_t stream);

/* out[r, :] (+)= sum_b x[b*batch_rows + row_offset + r, :]   (gradient of the broadcast above). accumulate != 0 adds to out. */
int gvk_batch_rowsum(const float* x, int ldx, int batch_rows, int row_offset, int R, int dim, int B, float* out, int accumulate, gvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * GAViKO gated prompt fusion — Awakening_Prompt between the two projections (model/gaviko.py:158-184, :20-47, :48-70, :84-119).
 * Inputs are the QuickGELU'd rank-r latents xl [B, T, r] (global: P prompts, cls, N image tokens) and ll [B, N, r] (local).
 *   imp  = sigmoid(W3 gelu(W1 LN_a(cl) + b1) + b3)            (PromptRelevantEstimator)     [B, P]
 *   gw   = sigmoid(Wg LN_g(cl) + bg)                          (PromptContextFusion)         [B]
 *   ctx_g[p] = softmax(r^-0.5 (Wqg pl[p] + bqg) . tok) tok,  tok = xl[:, 2P+2:]  (967 keys — the reference's double slice)
 *   ctx_l[p] = softmax(r^-0.5 (Wql pl[p] + bql) . tok) tok,  tok = ll
 *   enh[p] = (gw ctx_g[p] + (1 - gw) ctx_l[p]) * imp[p]   -> written IN PLACE over xl[:, p]  (xl then is `combined_latent`)
 * pl (the original prompt latents) and the small per-prompt state are saved for backward.
 * r in {16, 20, 32}; hidden width of the estimator fixed at 64 (model/gaviko.py:25).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* wq_g; const float* bq_g; const float* wq_l; const float* bq_l;   /* [r,r], [r] */
  const float* a_ln_w; const float* a_ln_b; const float* a_w1; const float* a_b1; const float* a_w3; const float* a_b3; /* [r],[r],[64,r],[64],[P,64],[P] */
  const float* g_ln_w; const float* g_ln_b; const float* g_w; const float* g_b;  /* [r],[r],[1,r],[1] */
} gvk_fusion_weights;

typedef struct {
  float* wq_g; float* bq_g; float* wq_l; float* bq_l;
  float* a_ln_w; float* a_ln_b; float* a_w1; float* a_b1; float* a_w3; float* a_b3;
  float* g_ln_w; float* g_ln_b; float* g_w; float* g_b;
} gvk_fusion_grads; /* accumulated with atomics: zero first */

typedef struct {
  float* pl;      /* [B, P, r]   original prompt latents */
  float* qg;      /* [B, P, r]   global queries */
  float* ql;      /* [B, P, r]   local queries */
  float* ctx_g;   /* [B, P, r] */
  float* ctx_l;   /* [B, P, r] */
  float* lse_g;   /* [B, P] */
  float* lse_l;   /* [B, P] */
  float* imp;     /* [B, P] */
  float* gw;      /* [B] */
} gvk_fusion_saved;

typedef struct {
  float* xl; const float* ll;
  int B, T, N, P, r;
  gvk_fusion_weights w;
  gvk_fusion_saved s;
} gvk_fusion_fwd_params;
int gvk_prompt_fusion_fwd(const gvk_fusion_fwd_params* p, gvk_stream_t stream);

/* Backward.  dxl holds dL/d(combined_latent) [B, T, r] on entry and dL/d(xl) on exit (in place); dll [B, N, r] is written.
 * ws: workspace of B*P*(2r + 4) floats. */
typedef struct {
  const float* xl;   /* combined latent as left by forward (rows >= P are the original xl rows) */
  const float* ll;
  float* dxl; float* dll;
  int B, T, N, P, r;
  gvk_fusion_weights w;
  gvk_fusion_saved s;
  gvk_fusion_grads g;
  float* ws;
} gvk_fusion_bwd_params;
int gvk_prompt_fusion_bwd(const gvk_fusion_bwd_params* p, gvk_stream_t stream);

/* K-extension operand of a bf16 GEMM for a rank-r fp32 product (bf16 mode: the Awakening_Prompt up-projection, model/gaviko.py:183-187, rides on
 * the fc2 GEMM of the same layer as 64 extra K columns).  dst[row, s*r + j] (bf16, s = 0..2) = hi(src[row, j]) or lo(src[row, j]) according to
 * bit s of `pattern` (0 = hi = bf16(x), 1 = lo = bf16(x - hi)); columns [3r, width) are zero.  Pack the A side with pattern 0b010 (hi, lo, hi)
 * and the B side with 0b100 (hi, hi, lo): the GEMM then accumulates hi*hi + lo*hi + hi*lo, the fp32 product to ~2^-16 relative. */
int gvk_split_pack_bf16(const float* src, int ld_src, int rows, int r, void* dst, int ld_dst, int width, int pattern, gvk_stream_t stream);

/* y = dy * quick_gelu'(pre)   over n elements (may run in place: y == dy). */
int gvk_quickgelu_bwd(const float* dy, const float* pre, float* y, size_t n, gvk_stream_t stream);
/* y = dy * (z > 0)   over n elements (ReLU backward from the activation output; Adapter, model/adaptformer.py:63; in place allowed). */
int gvk_relu_bwd(const float* dy, const float* z, float* y, size_t n, gvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * DVPT side path — share_MLP of model/dvpt.py:25-47:  prompt = (W_u cat[softmax(pl tok^T d_model^-0.5) tok ; cl ; tok] + b_u) * gate,
 * [pl ; cl ; tok] = W_d QuickGELU(x) + b_d.  The two projections are gvk_rowproj_down / gvk_rowproj_up on QuickGELU(x) and on gate-scaled
 * copies of (W_u, b_u); the pieces below are what DVPT adds.
 * ------------------------------------------------------------------------------------------------------------------ */
/* y = x * sigmoid(1.702 x) over n elements (n % 4 == 0): QuickGELU BEFORE the down-projection (model/dvpt.py:37). */
int gvk_quickgelu_fwd(const float* x, float* y, size_t n, gvk_stream_t stream);
/* y = res + dy * quick_gelu'(pre) over n elements (res optional; y may alias dy or res). */
int gvk_quickgelu_bwd_add(const float* dy, const float* pre, const float* res, float* y, size_t n, gvk_stream_t stream);

/* Cross attention of the P prompt latents over the N = T - P - 1 token latents of each volume (model/dvpt.py:38-45), r = 20:
 *   z[b, p] <- softmax(scale * z[b, p] . tok) tok,   tok = z[b, P+1:]          (IN PLACE over the prompt rows; cls and token rows untouched)
 * pl receives the prompt latents the rows held before (the queries), lse the log-sum-exp of every row of the attention. */
typedef struct {
  float* z; int B, T, P, r; float scale;
  float* pl;   /* [B, P, r] */
  float* lse;  /* [B, P] */
} gvk_latent_xattn_fwd_params;
int gvk_latent_xattn_fwd(const gvk_latent_xattn_fwd_params* p, gvk_stream_t stream);
/* Backward: dz holds d(combined latent) [B, T, r] on entry and d(latent before the attention) on exit (in place). */
typedef struct {
  const float* z;   /* combined latent as left by the forward */
  const float* pl; const float* lse;
  float* dz; int B, T, P, r; float scale;
} gvk_latent_xattn_bwd_params;
int gvk_latent_xattn_bwd(const gvk_latent_xattn_bwd_params* p, gvk_stream_t stream);

/* The scalar prompt_gate folded into the up-projection parameters (model/dvpt.py:30,46): y = gate[0] * x, and from the gradient dy of
 * y:  dx += gate[0] * dy,  dgate[0] += <x, dy>  (one CTA, deterministic). */
int gvk_gate_scale(const float* x, const float* gate, float* y, size_t n, gvk_stream_t stream);
int gvk_gate_grads(const float* x, const float* dy, const float* gate, float* dx, float* dgate, size_t n, gvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Head: final LayerNorm on the pooled rows only, mean-pool, Linear  (model/gaviko.py:306,314-316;
 * model/vision_transformer.py:159-164 with pool = 'cls' -> rows [0,1), 'mean' -> rows [0,T)).
 *   pooled[b] = mean_{r in [pool_start, pool_start+pool_count)} LN(x[b, r]) (* ssf_scale + ssf_shift);  logits = pooled Wh^T + bh
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* x; int ldx; int B, T, dim;
  int pool_start, pool_count;
  const float* gamma; const float* beta; float eps;
  const float* ssf_scale; const float* ssf_shift;
  const float* wh; const float* bh; int num_classes;
  float* pooled;   /* [B, dim] saved for backward */
  float* logits;   /* [B, num_classes] */
} gvk_head_fwd_params;
int gvk_head_fwd(const gvk_head_fwd_params* p, gvk_stream_t stream);

typedef struct {
  gvk_head_fwd_params f;
  const float* dlogits;     /* [B, num_classes] */
  float* dx; int ld_dx;     /* [B*T, dim]: only the pooled rows are written (zero the rest beforehand) */
  void* dx_lp; int ld_dx_lp; /* optional bf16 copy of the same rows */
  float* dwh; float* dbh;   /* [num_classes, dim], [num_classes]: overwritten, or accumulated into when accumulate_w != 0 */
  int accumulate_w;
  float* dgamma; float* dbeta;  /* optional [dim]: accumulated (atomics) */
  float* dssf_scale; float* dssf_shift; /* optional [dim]: accumulated (atomics) */
} gvk_head_bwd_params;
int gvk_head_bwd(const gvk_head_bwd_params* p, gvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * --method evp (reference src/model/evp.py): the two primitives beyond the frozen-ViT kernel set.
 *
 * gvk_wgrad: general weight gradient   dw[i, j] += sum_{m < M} a[row_a(m), i] * b[row_b(m), j]      (i < na, j < nb; dw fp32, ACCUMULATES)
 *   row_x(m) = (m / x_rows_per_batch) * x_batch_rows + m % x_rows_per_batch   (x_rows_per_batch == 0: identity) — lets an operand skip the
 *   cls row of every volume without a copy.  a / b are fp32 or bf16, row-major, na / nb / lda / ldb multiples of 4.
 *   precision GVK_PREC_TF32: mma.sync tf32 tensor-core tiles;  GVK_PREC_FP32: exact FFMAs.  Rows are split over CTAs and combined with
 *   fp32 atomics (run-to-run differences at the last bit).  Replaces autograd's weight gradients of PromptGenerator.shared_mlp,
 *   embedding_generator, lightweight_mlp_i and prompt_generator.proj (model/evp.py:42-52), whose rank dim / scale_factor (192 at ViT-B
 *   with configs/evp.yaml) is past the r <= 32 limit of gvk_skinny_wgrad.
 *
 * gvk_hfreq_filter: PromptGenerator.fft (model/evp.py:124-146) in closed form.  For a (B, C, D, H, W) volume the reference's
 *   fft2 (last two axes) + all-axes fftshift + mask[:, :, a:b, c:d] = 1 (axes D and H!) + inverse + real + abs equals: on the depth slices
 *   with hit[d] != 0, out = | filt @ slice |  (filt = I - Re(IDFT diag(cut) DFT), an H x H symmetric real matrix acting along H);
 *   on the others out = | slice |.  in / out: fp32 [slices = B*C*D, H, W]; filt fp32 [H, H]; hit uint8 [D].  Exact fp32 FMAs.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a; int a_dtype; int lda; int na; int a_rows_per_batch, a_batch_rows;
  const void* b; int b_dtype; int ldb; int nb; int b_rows_per_batch, b_batch_rows;
  int M;
  float* dw; int ld_dw;
  int precision;
} gvk_wgrad_params;
int gvk_wgrad(const gvk_wgrad_params* p, gvk_stream_t stream);

typedef struct {
  const float* in; float* out;
  const float* filt;
  const unsigned char* hit;
  int slices, D, H, W;
} gvk_hfreq_filter_params;
int gvk_hfreq_filter(const gvk_hfreq_filter_params* p, gvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Losses on [B, C] fp32 logits, int64 targets.  loss is a device scalar; dlogits (optional) is d loss / d logits.
 * kind 0: the reference's FocalLoss (losses/focal_loss.py:84-111: clamp(logits) -> softmax -> clamp -> softmax, eps 1e-16,
 *         ignore_index, mean over non-ignored);  kind 1: nn.CrossEntropyLoss (mean).
 * ------------------------------------------------------------------------------------------------------------------ */
int gvk_loss_fwd_bwd(const float* logits, const long long* target, int B, int C, int kind, float gamma, float eps, long long ignore_index,
                     float* loss, float* dlogits, gvk_stream_t stream);


/* ------------------------------------------------------------------------------------------------------------------
 * Optimiser step on the flat trainable buffer (reference src/train.py:315-319: clip_grad_norm_(params, 1.0) then Adam.step()).
 *   gvk_grad_sumsq : partials[i] = sum of squares of a slice of (grad * grad_scale); n_partials = GVK_SUMSQ_PARTIALS
 *   gvk_clip_adam  : norm = sqrt(sum partials); coef = min(1, max_norm / (norm + 1e-6))  (max_norm <= 0 disables clipping);
 *                    g = grad * grad_scale * coef + weight_decay * p;  torch.optim.Adam update with bias correction at `step` (1-based).
 * grad_scale carries the 1/world_size of the data-parallel mean, so the clip sees the gradient of the GLOBAL batch.
 * Both are deterministic (fixed partition, no atomics): every rank computes bit-identical updates from the all-reduced buffer.
 * ------------------------------------------------------------------------------------------------------------------ */
enum { GVK_SUMSQ_PARTIALS = 128 };
int gvk_grad_sumsq(const float* grad, size_t n, float grad_scale, float* partials, gvk_stream_t stream);
int gvk_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, const float* partials, float max_norm, float grad_scale,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step, float* grad_norm_out, gvk_stream_t stream);
/* gvk_clip_adam with the learning rate (float) and the 1-based step (int64) read from DEVICE memory: the form a captured CUDA graph replays
 * (gaviko_b200/graph.py), where both change between replays of the same kernel node. */
int gvk_clip_adam_dyn(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, const float* partials, float max_norm, float grad_scale,
                      const float* lr_dev, float beta1, float beta2, float eps, float wd, const long long* step_dev, float* norm_out, gvk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GVK_H_ */
