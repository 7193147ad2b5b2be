// gvk_patch_embed.cu — Conv3d patch embedding (kernel = stride = patch) as ONE kernel: TMA gathers the non-overlapping patches of the fp32
// volume straight into the operand ring of a tcgen05 GEMM (kind::tf32, fp32 accumulators in TMEM); the epilogue adds bias + positional
// embedding and writes the token rows into both token streams.  No im2col matrix, no separate gather pass, no transpose copy.
//
// Geometry.  volume (B, C, D, H, W) fp32, patch (fp, ps, ps), grid (gd, gh, gw); token m = ((b gd + d) gh + h) gw + w, weight column
// k = ((c fp + kd) ps + kh) ps + kw (model/gaviko.py:383-385,532-533).  A "plane" is the gh x gw tokens of one (b, d); a K block is one
// (c, kd, kh): the ps = 16 consecutive kw of every token, one 64-byte segment of the volume per token.  The whole [gh gw tokens x 16] block
// of a plane is ONE 5-D TMA box
//     dims (kw, kh, w, h, z = (b C + c) D + d fp + kd),  box (16, 1, gw, gh, 1)
// which lands in shared memory token-major, SWIZZLE_128B.  TMA gives every inner row its own 128-byte swizzle line (measured: a 64-byte
// inner box is NOT packed two to a line), so a line holds 16 floats and the tile is a K-major UMMA operand of which two K = 8 steps are
// used; the weight tile is loaded with the same 16-float inner box and has the same pitch.
//
// The tokens are the N side of the MMA (N = planes_per_tile x tokens-per-plane rounded to 8, <= 256; 2 x 104 = 208 for the 10 x 10 x 10
// grid), the embedding dimension is the M side (128 weight rows per tile, fp32 weight exactly as nn.Conv3d stores it, 2-D TMA).  The
// accumulator therefore has lane = output feature, column = token: each epilogue store of a warp is 32 consecutive features of one token
// row (128 contiguous bytes) with no shared-memory transpose.  tf32 keeps 10 mantissa bits of the fp32 volume / weight (bf16 keeps 7).
//
// Replaces Gaviko.conv_proj + flatten/transpose + the token assembly of the patch rows (model/gaviko.py:383-385,532-548).
#include <algorithm>

#include "gvk_common.cuh"

namespace gvk {

namespace pe {
constexpr int kThreads = 384;          // warp 0 TMA, warp 1 MMA, warp 2 TMEM, warps 4-11 epilogue
constexpr int kBM = 128;               // output features per tile
constexpr int kBK = 16;                // floats per K block (64 of the 128 bytes of a swizzle line)
constexpr int kLine = 128;             // bytes per operand row in shared memory
constexpr int kStages = 4;
constexpr int kABytes = kBM * kLine;        // 16 KB
constexpr int kBBytesMax = 256 * kLine;    // 32 KB
constexpr int kATx = kBM * kBK * 4;        // bytes the weight box delivers
constexpr int kSmem = kStages * (kABytes + kBBytesMax) + (2 * kStages + 4) * 8 + 16 + 1024;

struct Args {
  int dim, n_tok, tok_plane, rows_pad, planes_per_tile, n_planes, n_kb, kb_per_z, gd, fp, C, D;
  int n_dim_tiles, n_tiles;
  const float* bias;
  const float* pos;
  float* out;
  int ld_out, out_batch_rows, out_row_offset;
  float* out2;
  int ld_out2;
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
patch_embed_tf32_kernel(const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_img, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // [kStages][128 features x 32 floats]
  uint8_t* sB = smem + kStages * kABytes;               // [kStages][up to 256 tokens x 32 floats]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + kStages * kBBytesMax);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_img);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int box_tx = a.tok_plane * kBK * 4;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: per K block one weight box + one box per plane of the tile =====
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        const int pg = t / a.n_dim_tiles, nt = t - pg * a.n_dim_tiles;      // consecutive CTAs share the plane group (its patches stay in L2)
        const int plane0 = pg * a.planes_per_tile;
        const int np = min(a.planes_per_tile, a.n_planes - plane0);
        for (int kb = 0; kb < a.n_kb; ++kb) {
          const int zk = kb / a.kb_per_z, tk = kb - zk * a.kb_per_z;         // zk = c * fp + kd, tk = kh
          const int c = zk / a.fp, kd = zk - c * a.fp;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kATx + np * box_tx);
          tma_load_2d(sA + stage * kABytes, &tma_w, &full_bar[stage], kb * kBK, nt * kBM);
          for (int p = 0; p < np; ++p) {
            const int plane = plane0 + p;
            const int b = plane / a.gd, d = plane - b * a.gd;
            tma_load_5d(sB + stage * kBBytesMax + p * a.rows_pad * kLine, &tma_img, &full_bar[stage], 0, tk, 0, 0, (b * a.C + c) * a.D + d * a.fp + kd);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform control flow, tcgen05 instructions under elect_one) =====
    const uint32_t idesc = make_idesc_tf32(kBM, a.planes_per_tile * a.rows_pad);
    int stage = 0;
    uint32_t phase = 0;
    int lt = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++lt) {
      const int as = lt & 1;
      mbar_wait(&tempty_bar[as], ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 256;
      for (int kb = 0; kb < a.n_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(sA + stage * kABytes), b_addr = smem_u32(sB + stage * kBBytesMax);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k)      // 8 tf32 = 32 bytes per K step
            umma_tf32(d_tmem, make_sw128_desc(a_addr + k * 32, 16, 1024), make_sw128_desc(b_addr + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == a.n_kb - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: lane = output feature, column = token; + bias + positional embedding -> both token streams =====
    const int ew = warp - 4, q = warp & 3, half = ew >> 2;
    const int ncols = a.planes_per_tile * a.rows_pad;
    const int nchunks = (ncols + 31) / 32;
    int lt = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++lt) {
      const int pg = t / a.n_dim_tiles, nt = t - pg * a.n_dim_tiles;
      const int plane0 = pg * a.planes_per_tile;
      const int as = lt & 1;
      const int n = nt * kBM + q * 32 + lane;
      const bool nvalid = n < a.dim;
      const float bias = (nvalid && a.bias) ? a.bias[n] : 0.f;
      mbar_wait(&tfull_bar[as], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
      for (int ch = half; ch < nchunks; ch += 2) {
        float v[32];
        tmem_ld_32x32(taddr0 + ch * 32, v);
        tc_wait_ld();
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const int col = ch * 32 + j;
          const int p = col / a.rows_pad, r = col - p * a.rows_pad;
          const int plane = plane0 + p;
          if (col < ncols && r < a.tok_plane && plane < a.n_planes && nvalid) {      // warp-uniform except nvalid
            const int m = plane * a.tok_plane + r;
            const int b = m / a.n_tok, tok = m - b * a.n_tok;
            float x = v[j] + bias;
            if (a.pos) x += a.pos[(size_t)tok * a.dim + n];
            a.out[((size_t)b * a.out_batch_rows + a.out_row_offset + tok) * a.ld_out + n] = x;
            if (a.out2) a.out2[(size_t)m * a.ld_out2 + n] = x;
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
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}
}  // namespace pe

// Geometry the kernel covers; everything else goes through gvk_patch_gather + gvk_gemm.
static bool patch_embed_geometry(const gvk_patch_embed_params* p, pe::Args* a) {
  if (p->ps != 16 || p->fp <= 0 || p->C <= 0 || p->D % p->fp || p->H % p->ps || p->W % p->ps) return false;
  const int gd = p->D / p->fp, gh = p->H / p->ps, gw = p->W / p->ps;
  const int tok_plane = gh * gw;
  if (gh > 256 || gw > 256 || tok_plane > 256) return false;
  const int rows_pad = (tok_plane + 7) / 8 * 8;
  int ppt = 256 / rows_pad;
  while (ppt > 1 && (ppt * rows_pad) % 16 != 0) --ppt;
  if ((ppt * rows_pad) % 16 != 0 || ppt * rows_pad < 16) return false;
  if ((p->W * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(p->img) & 15) || (reinterpret_cast<uintptr_t>(p->weight) & 15)) return false;
  if (a) {
    a->dim = p->dim; a->n_tok = gd * tok_plane; a->tok_plane = tok_plane; a->rows_pad = rows_pad; a->planes_per_tile = ppt;
    a->n_planes = p->B * gd; a->kb_per_z = p->ps; a->n_kb = p->C * p->fp * p->ps; a->gd = gd; a->fp = p->fp; a->C = p->C; a->D = p->D;
    a->n_dim_tiles = (p->dim + pe::kBM - 1) / pe::kBM;
    a->n_tiles = a->n_dim_tiles * ((a->n_planes + ppt - 1) / ppt);
  }
  return true;
}

int patch_embed_supported(const gvk_patch_embed_params* p) { return p && patch_embed_geometry(p, nullptr) ? 1 : 0; }

int patch_embed(const gvk_patch_embed_params* p, cudaStream_t stream) {
  GVK_CHECK_ARG(p && p->img && p->weight && p->out, "gvk_patch_embed: null pointer");
  GVK_CHECK_ARG(p->B > 0 && p->dim > 0, "gvk_patch_embed: non-positive shape");
  pe::Args a;
  GVK_CHECK_ARG(patch_embed_geometry(p, &a), "gvk_patch_embed: unsupported geometry (needs ps = 16, <= 256 tokens per plane); use gvk_patch_gather + gvk_gemm");
  a.bias = p->bias; a.pos = p->pos; a.out = p->out; a.ld_out = p->ld_out; a.out_batch_rows = p->out_batch_rows; a.out_row_offset = p->out_row_offset;
  a.out2 = p->out2; a.ld_out2 = p->ld_out2;
  static bool configured = false;
  if (!configured) {
    int st = cuda_status(cudaFuncSetAttribute(pe::patch_embed_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pe::kSmem), "patch_embed smem");
    if (st != GVK_OK) return st;
    configured = true;
  }
  CUtensorMap tw, ti;
  const int K = p->C * p->fp * p->ps * p->ps;
  const uint64_t wdims[2] = {(uint64_t)K, (uint64_t)p->dim}, wstrides[1] = {(uint64_t)K * 4};
  const uint32_t wbox[2] = {pe::kBK, pe::kBM};
  int st = make_tma_f32(&tw, p->weight, 2, wdims, wstrides, wbox);
  if (st != GVK_OK) return st;
  const int gh = p->H / p->ps, gw = p->W / p->ps;
  const uint64_t dims[5] = {(uint64_t)p->ps, (uint64_t)p->ps, (uint64_t)gw, (uint64_t)gh, (uint64_t)p->B * p->C * p->D};
  const uint64_t strides[4] = {(uint64_t)p->W * 4, (uint64_t)p->ps * 4, (uint64_t)p->ps * p->W * 4, (uint64_t)p->H * p->W * 4};
  const uint32_t box[5] = {(uint32_t)p->ps, 1u, (uint32_t)gw, (uint32_t)gh, 1u};
  st = make_tma_f32(&ti, p->img, 5, dims, strides, box);
  if (st != GVK_OK) return st;
  const int grid = std::min(a.n_tiles, sm_count());
  pe::patch_embed_tf32_kernel<<<grid, pe::kThreads, pe::kSmem, stream>>>(tw, ti, a);
  GVK_CHECK_LAUNCH("patch_embed_tf32");
  return GVK_OK;
}

}  // namespace gvk
