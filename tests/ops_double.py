"""TEST DOUBLE of ``gaviko_b200.ops``: every wrapper restated in plain torch fp32 on the CPU, contract by contract from include/gvk.h.

Purpose: exercise the HOST logic of the engines (``engine.py``, ``vit_engine.py``, ``dvpt_engine.py``: call order, in-place conventions, token /
row bookkeeping, which gradients go where) in the ``-m "not gpu"`` suite, against the same golden vectors the GPU tests use.  It is test
infrastructure only: it is installed by monkeypatching the engines' ``ops`` reference (``install()`` below) and never ships — the product path has no
CPU fallback (``tests/test_dropin_surface.py::test_no_cpu_fallback``).  The arithmetic inside every wrapper is torch fp32 whatever the engine's
compute mode (bf16 mode: bf16 storage of operands / activations, the hi / lo slot packing and the one-pass forms as compositions of the kernels
they replace); the one-kernel patch embedding reports 'not taken'.  Dropout masks are replayable functions of the seed (not the kernels' Philox
values); attention-probability dropout is not emulated.
"""
import contextlib
import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX = 0, 1, 2, 3, 4
ROWACT_NONE, ROWACT_QUICKGELU, ROWACT_RELU = 0, 1, 2
LOSS_FOCAL, LOSS_CE = 0, 1
PREC_FP32, PREC_TF32 = 0, 1
GEMM_HOOK = None


class GvkError(RuntimeError):
    pass


def _keep(shape, drop_p, seed, offset=0):
    """Replayable dropout multiplier (0 or 1 / (1 - p)) of a tensor: a function of (seed, offset, shape) only, like the Philox masks of the
    kernels (the VALUES differ from the CUDA masks; forward / backward consistency is what the engines rely on)."""
    if drop_p <= 0.0:
        return None
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(offset)) % (2 ** 63 - 1))
    return (torch.rand(shape, generator=g) >= drop_p).float() / (1.0 - drop_p)


def _gelu_grad(x):
    return 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * 0.3989422804014327 * torch.exp(-0.5 * x * x)


def _qg(x):
    return x * torch.sigmoid(1.702 * x)


def _qg_grad(x):
    s = torch.sigmoid(1.702 * x)
    return s * (1 + 1.702 * x * (1 - s))


# ---------------------------------------------------------------------------------------------- GEMM
def gemm(a, b, *, out=None, out_dtype=None, bias=None, ssf_scale=None, ssf_shift=None, act=ACT_NONE, aux=None, pos=None, rows_per_batch=0,
         out_batch_rows=0, out_row_offset=0, res1=None, res2=None, out2=None, out_rows=None):
    M, N = a.shape[0], b.shape[0]
    v = a.float() @ b.float().t()
    if bias is not None:
        v = v + bias
    if ssf_scale is not None:
        v = v * ssf_scale + ssf_shift
    if act == ACT_GELU:
        if aux is not None:
            aux.copy_(v)
        v = F.gelu(v)
    elif act == ACT_GELU_BWD:
        v = v * _gelu_grad(aux.float())
    elif act == ACT_GELU_SAVE_GRAD:
        if aux is not None:
            aux.copy_(_gelu_grad(v))
        v = F.gelu(v)
    elif act == ACT_MUL_AUX:
        v = v * aux.float()
    rows = torch.arange(M)
    if rows_per_batch > 0:
        bidx, r = rows // rows_per_batch, rows % rows_per_batch
        if pos is not None:
            v = v + pos[r]
        rows = bidx * out_batch_rows + out_row_offset + r
    if res1 is not None:
        v = v + res1
    if res2 is not None:
        v = v + res2
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, N), dtype=out_dtype or torch.float32)
    out[rows] = v.to(out.dtype)
    if out2 is not None:
        out2.copy_(v)
    return out


# ---------------------------------------------------------------------------------------------- row kernels
def patch_embed(*args, **kwargs):
    """The one-kernel TMA gather + tf32 GEMM has no restatement of its own: report 'not taken' so the engine uses patch_gather + gemm."""
    return False


def layernorm_fwd(x, gamma, beta, *, out=None, out_dtype=torch.float32, eps=1e-5, ssf_scale=None, ssf_shift=None, save_stats=True):
    mean = x.mean(1)
    rstd = torch.rsqrt(x.var(1, unbiased=False) + eps)
    y = (x - mean[:, None]) * rstd[:, None] * gamma + beta
    if ssf_scale is not None:
        y = y * ssf_scale + ssf_shift
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype)
    out.copy_(y)
    return out, (mean if save_stats else None), (rstd if save_stats else None)


def layernorm_fwd_down(x, gamma, beta, w, bias=None, *, act=ROWACT_NONE, save_pre=False, eps=1e-5, save_stats=True):
    y, mean, rstd = layernorm_fwd(x, gamma, beta, out_dtype=torch.bfloat16, eps=eps, save_stats=save_stats)
    d = rowproj_down(x, w, bias, act=act, save_pre=save_pre)
    return y, mean, rstd, dict(z=d['z'], pre=d['pre'])


def layernorm_fwd_down_supported(x, r):
    return x.shape[1] in (384, 768) and r <= 32


def rowproj_down(x, w, bias=None, *, transposed=False, ln=None, eps=1e-5, act=ROWACT_NONE, save_pre=False, w2=None, drop_p=0.0, seed=0, offset=0, prec=PREC_FP32):
    mean = rstd = z2 = None
    k = _keep(x.shape, drop_p, seed, offset)
    fx = x if k is None else x * k
    x = fx
    if ln is not None:
        mean = x.mean(1)
        rstd = torch.rsqrt(x.var(1, unbiased=False) + eps)
        fx = (x - mean[:, None]) * rstd[:, None] * ln[0] + ln[1]
    W = w.t() if transposed else w            # [r, dim]
    pre = fx @ W.t()
    if bias is not None:
        pre = pre + bias
    z = _qg(pre) if act == ROWACT_QUICKGELU else (torch.relu(pre) if act == ROWACT_RELU else pre)
    z = z.contiguous().clone()
    if w2 is not None:
        z2 = z @ w2.t()
    return dict(z=z, pre=pre.clone() if save_pre else None, z2=z2, mean=mean, rstd=rstd)


def rowproj_up(c, w, bias=None, *, transposed=False, res=None, out=None, out_lp=None, drop_p=0.0, seed=0, offset=0, prec=PREC_FP32):
    v = c @ (w if transposed else w.t())
    if bias is not None:
        v = v + bias
    k = _keep(v.shape, drop_p, seed, offset)
    if k is not None:
        v = v * k
    if res is not None:
        v = v + res
    if out is None:
        out = torch.empty_like(v)
    out.copy_(v)
    if out_lp is not None:
        out_lp.copy_(v)
    return out


def rowproj_up_down(c, w, bias=None, *, transposed=False, res=None, out=None, up_drop_p=0.0, up_seed=0, up_offset=0, w2=None, bias2=None, transposed2=False,
                    act=ROWACT_NONE, save_pre=False, dn_drop_p=0.0, dn_seed=0, dn_offset=0):
    out = rowproj_up(c, w, bias, transposed=transposed, res=res, out=out, drop_p=up_drop_p, seed=up_seed, offset=up_offset)
    d = rowproj_down(out, w2, bias2, transposed=transposed2, act=act, save_pre=save_pre, drop_p=dn_drop_p, seed=dn_seed, offset=dn_offset)
    return out, dict(z=d['z'], pre=d['pre'])


def rowproj_up_down_supported(dim, r, r2, prec):
    return prec == PREC_TF32 and dim in (384, 768) and r <= 24 and r2 <= 24


def skinny_wgrad(a, x, *, dw=None, dw_layout='rd', da_colsum=None, dx_colsum=None, ln=None, drop_p=0.0, seed=0, offset=0, prec=PREC_FP32, dw_strides=None):
    k = _keep(x.shape, drop_p, seed, offset)
    fx = x if k is None else x * k
    x = fx
    if ln is not None:
        fx = (x - ln[2][:, None]) * ln[3][:, None] * ln[0] + ln[1]
    if dw is not None:
        if dw_strides is not None:
            assert dw_strides[0] == 1          # element (j, c) at c * sc + j: a [dim, r] view of a wider gradient
            dw.add_(fx.t() @ a)
        elif dw_layout == 'rd':
            dw.add_(a.t() @ fx)
        else:
            dw.add_(fx.t() @ a)
    if da_colsum is not None:
        da_colsum.add_(a.sum(0))
    if dx_colsum is not None:
        dx_colsum.add_(fx.sum(0))


def layernorm_bwd(x, gamma, mean, rstd, *, dy=None, dz=None, w=None, dres=None, dx=None, dx_lp=None, dgamma=None, dbeta=None, az=None, aw=None,
                  beta=None, ssf_scale=None, dssf_scale=None, dssf_shift=None, prec=0, ow=None, ow_transposed=False):
    xh = (x - mean[:, None]) * rstd[:, None]
    g = torch.zeros_like(x)
    if dy is not None:
        g = g + dy
    if dz is not None:
        g = g + dz @ w
    if ssf_scale is not None:
        y_ln = xh * gamma + beta
        if dssf_scale is not None:
            dssf_scale.add_((g * y_ln).sum(0))
        if dssf_shift is not None:
            dssf_shift.add_(g.sum(0))
        g = g * ssf_scale
    if dgamma is not None:
        dgamma.add_((g * xh).sum(0))
    if dbeta is not None:
        dbeta.add_(g.sum(0))
    gx = g * gamma
    v = rstd[:, None] * (gx - gx.mean(1, keepdim=True) - xh * (gx * xh).mean(1, keepdim=True))
    if dres is not None:
        v = v + dres
    if az is not None:
        v = v + az @ aw
    if dx is None:
        dx = torch.empty_like(x)
    dx.copy_(v)
    if dx_lp is not None:
        dx_lp.copy_(v)
    if ow is not None:
        return dx, (v @ (ow if ow_transposed else ow.t())).contiguous()
    return dx


def layernorm_bwd_down_supported(x, orank, prec):
    return prec == PREC_TF32 and x.shape[1] in (384, 768) and orank <= 24


def small_wgrad(a, b, dw):
    dw.add_(a.t() @ b)


def small_matmul(a, w):
    return a @ w


def colsum(x, out):
    out.add_(x.sum(0))


def cast_bf16(x, out=None):
    return x.clone() if out is None else out.copy_(x)       # fp32-only double


def cast_f32(x, out=None):
    return x.float().clone() if out is None else out.copy_(x)


def ssf_bwd(dy, *, y=None, scale=None, shift=None, dx=None, dscale=None, dshift=None, sub=None, rows_per_batch=0, batch_rows=0, M=None):
    M = dy.shape[0] if M is None else M
    m = torch.arange(M)
    rows = (m // rows_per_batch) * batch_rows + m % rows_per_batch if rows_per_batch > 0 else m
    d = dy[rows].float()
    if dshift is not None:
        dshift.add_(d.sum(0))
    if scale is not None:
        xin = y[rows].float()
        if sub is not None:
            xin = xin - sub[m % rows_per_batch]
        xin = (xin - shift) / scale
        if dscale is not None:
            dscale.add_((d * xin).sum(0))
        if dx is not None:
            dx[rows] = (d * scale).to(dx.dtype)
    return dx


def dropout(x, drop_p, seed, *, res=None, out=None, out_dtype=None, offset=0):
    k = _keep(x.shape, drop_p, seed, offset)
    v = (x.float() if k is None else x.float() * k) + (res if res is not None else 0)
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype or x.dtype)
    return out.copy_(v)


def relu_bwd(dy, z, out=None):
    v = dy * (z > 0)
    return v if out is None else out.copy_(v)


def quickgelu_bwd(dy, pre, out=None):
    v = dy * _qg_grad(pre)
    return v if out is None else out.copy_(v)


# ---------------------------------------------------------------------------------------------- attention
def _window_allow(window, grid):
    D, H, W = grid
    idx = torch.arange(D * H * W)
    coords = torch.stack([idx // (H * W), (idx // W) % H, idx % W], 1)
    allow = torch.ones(D * H * W, D * H * W, dtype=torch.bool)
    for ax, k in enumerate(window):
        ci, cj = coords[:, ax][:, None], coords[:, ax][None, :]
        allow &= (cj >= ci - k // 2) & (cj <= ci + k - 1 - k // 2)
    return allow


def _split_heads(qkv, B, T, H, D, off):
    return qkv[:, off:off + H * D].float().reshape(B, T, H, D).transpose(1, 2)      # B H T D


def _attn(qkv, B, T, H, D, q_off, k_off, v_off, scale, window, grid):
    q, k, v = (_split_heads(qkv, B, T, H, D, o) for o in (q_off, k_off, v_off))
    s = q @ k.transpose(-1, -2) * scale
    if window is not None:
        s = s.masked_fill(~_window_allow(window, grid), float('-inf'))
    lse = torch.logsumexp(s, -1)
    p = torch.exp(s - lse[..., None])
    return q, k, v, p, lse


def attn_simt_fwd(qkv, B, T, H, D, *, q_off, k_off, v_off, scale, window=None, grid=None, drop_p=0.0, seed=0, offset=0, prec=PREC_FP32):
    q, k, v, p, lse = _attn(qkv, B, T, H, D, q_off, k_off, v_off, scale, window, grid)
    keep = _keep(p.shape, drop_p, seed, offset)
    out = ((p if keep is None else p * keep) @ v).transpose(1, 2).reshape(B * T, H * D).to(qkv.dtype)
    return out, lse.reshape(-1)


def attn_simt_bwd(qkv, out, lse, dout, B, T, H, D, *, q_off, k_off, v_off, scale, window=None, grid=None, drop_p=0.0, seed=0, offset=0, dqkv=None,
                  prec=PREC_FP32):
    q, k, v, p, _ = _attn(qkv, B, T, H, D, q_off, k_off, v_off, scale, window, grid)
    do = dout.float().reshape(B, T, H, D).transpose(1, 2)
    o = out.float().reshape(B, T, H, D).transpose(1, 2)
    keep = _keep(p.shape, drop_p, seed, offset)
    dp = do @ v.transpose(-1, -2)
    if keep is not None:
        dp = dp * keep
    ds = p * (dp - (do * o).sum(-1, keepdim=True))
    if keep is not None:
        p = p * keep                     # dV uses the dropped probabilities
    if dqkv is None:
        dqkv = torch.zeros_like(qkv)
    for off, t in ((q_off, ds @ k * scale), (k_off, ds.transpose(-1, -2) @ q * scale), (v_off, p.transpose(-1, -2) @ do)):
        dqkv[:, off:off + H * D] = t.transpose(1, 2).reshape(B * T, H * D).to(dqkv.dtype)
    return dqkv


def mhsa_fwd(qkv, B, T, H, scale, drop_p=0.0, seed=0):
    assert drop_p == 0.0, 'ops double: attention dropout is not emulated (the parity cases run with dropout off)'
    out, lse = attn_simt_fwd(qkv.float(), B, T, H, 64, q_off=0, k_off=H * 64, v_off=2 * H * 64, scale=scale)
    return out.to(qkv.dtype), lse


def mhsa_bwd(qkv, out, lse, dout, B, T, H, scale, drop_p=0.0, seed=0):
    assert drop_p == 0.0, 'ops double: attention dropout is not emulated (the parity cases run with dropout off)'
    return attn_simt_bwd(qkv.float(), out.float(), lse, dout.float(), B, T, H, 64, q_off=0, k_off=H * 64, v_off=2 * H * 64, scale=scale).to(qkv.dtype)


# ---------------------------------------------------------------------------------------------- token assembly
def rescale_intensity(x, out_min=0.0, out_max=1.0, *, out=None, out_dtype=None):
    B = x.shape[0]
    f = x.reshape(B, -1)
    lo, hi = f.min(1).values, f.max(1).values
    rng = (hi - lo).reshape([B] + [1] * (x.dim() - 1))
    v = (x - lo.reshape(rng.shape)) / rng * (out_max - out_min) + out_min
    return v.to(out_dtype or torch.float32) if out is None else out.copy_(v)


def split_pack_bf16(src, dst, pattern):
    """hi / lo bf16 slots of gvk_split_pack_bf16 (three r-wide slots, zero up to the width of dst)."""
    rows, r = src.shape
    hi = src.to(torch.bfloat16)
    lo = (src - hi.float()).to(torch.bfloat16)
    dst.zero_()
    for slot in range(3):
        dst[:, slot * r:(slot + 1) * r] = lo if (pattern >> slot) & 1 else hi
    return dst


def patch_gather(img, fp, ps, out_dtype):
    B, C, D, H, W = img.shape
    nd, nh, nw = D // fp, H // ps, W // ps
    x = img.reshape(B, C, nd, fp, nh, ps, nw, ps).permute(0, 2, 4, 6, 1, 3, 5, 7)
    return x.reshape(B * nd * nh * nw, C * fp * ps * ps).to(out_dtype).contiguous()


def _rowmap(M, rows):
    m = torch.arange(M)
    return m if not rows or rows[0] <= 0 else (m // rows[0]) * rows[1] + m % rows[0]


def wgrad(a, b, dw, *, M=None, a_rows=None, b_rows=None, prec=PREC_FP32):
    M = a.shape[0] if M is None else M
    dw.add_(a[_rowmap(M, a_rows)].float().t() @ b[_rowmap(M, b_rows)].float())
    return dw


def hfreq_filter(img, filt, hit):
    y = img.clone()
    sel = hit.bool()
    y[:, :, sel] = torch.einsum('hk,bcdkw->bcdhw', filt, img[:, :, sel])
    return y.abs()


def fill_rows(a, b, out, out_batch_rows, out_row_offset, B):
    R = a.shape[0]
    v = a + (b if b is not None else 0)
    for i in range(B):
        out[i * out_batch_rows + out_row_offset:i * out_batch_rows + out_row_offset + R] = v


def batch_rowsum(x, batch_rows, row_offset, R, B, out=None, accumulate=False):
    s = x.reshape(B, batch_rows, -1)[:, row_offset:row_offset + R].sum(0)
    if out is None:
        return s.clone()
    if accumulate:
        out.add_(s.reshape(out.shape))
    else:
        out.copy_(s.reshape(out.shape))
    return out


# ---------------------------------------------------------------------------------------------- prompt fusion (model/gaviko.py:158-184)
FUSION_WEIGHT_FIELDS = ('wq_g', 'bq_g', 'wq_l', 'bq_l', 'a_ln_w', 'a_ln_b', 'a_w1', 'a_b1', 'a_w3', 'a_b3', 'g_ln_w', 'g_ln_b', 'g_w', 'g_b')


def _fusion(xl3, ll3, w, P):
    r = xl3.shape[-1]
    pl, cl, il = xl3[:, :P], xl3[:, P:P + 1], xl3[:, P + 1:]
    h = F.layer_norm(cl, (r,), w['a_ln_w'], w['a_ln_b'])
    imp = torch.sigmoid(F.gelu(h @ w['a_w1'].t() + w['a_b1']) @ w['a_w3'].t() + w['a_b3'])
    gw = torch.sigmoid(F.layer_norm(cl, (r,), w['g_ln_w'], w['g_ln_b']) @ w['g_w'].t() + w['g_b'])

    def xattn(tok, wq, bq):
        a = torch.softmax((pl @ wq.t() + bq) @ tok.transpose(-1, -2) * r ** -0.5, -1)
        return a @ tok
    fused = gw * xattn(il[:, P + 1:], w['wq_g'], w['bq_g']) + (1 - gw) * xattn(ll3, w['wq_l'], w['bq_l'])
    return fused * imp.transpose(1, 2)


def prompt_fusion_fwd(xl, ll, w, B, T, N, P):
    r = xl.shape[1]
    saved = dict(xl0=xl.clone())
    enh = _fusion(xl.reshape(B, T, r), ll.reshape(B, N, r), w, P)
    xl.reshape(B, T, r)[:, :P] = enh
    return saved


def prompt_fusion_bwd(xl, ll, dxl, w, saved, grads, B, T, N, P):
    r = xl.shape[1]
    x0 = saved['xl0'].clone().requires_grad_(True)
    l0 = ll.clone().requires_grad_(True)
    ws = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    with torch.enable_grad():
        x3 = x0.reshape(B, T, r)
        comb = torch.cat([_fusion(x3, l0.reshape(B, N, r), ws, P), x3[:, P:]], 1)
        gs = torch.autograd.grad(comb, [x0, l0] + [ws[k] for k in FUSION_WEIGHT_FIELDS], dxl.reshape(B, T, r), allow_unused=True)
    dxl.copy_(gs[0])
    for k, g in zip(FUSION_WEIGHT_FIELDS, gs[2:]):
        if g is not None:
            grads[k].add_(g)
    return gs[1].contiguous()


# ---------------------------------------------------------------------------------------------- DVPT side path (model/dvpt.py:36-47)
def quickgelu_fwd(x, out=None):
    return _qg(x) if out is None else out.copy_(_qg(x))


def quickgelu_bwd_add(dy, pre, res=None, out=None):
    v = dy * _qg_grad(pre) + (res if res is not None else 0)
    return v if out is None else out.copy_(v)


def latent_xattn_fwd(z, B, T, P, scale):
    r = z.shape[1]
    z3 = z.reshape(B, T, r)
    pl, tok = z3[:, :P].clone(), z3[:, P + 1:]
    s = pl @ tok.transpose(-1, -2) * scale
    lse = torch.logsumexp(s, -1)
    z3[:, :P] = torch.exp(s - lse[..., None]) @ tok
    return pl.reshape(B * P, r), lse.reshape(-1)


def latent_xattn_bwd(z, pl, lse, dz, B, T, P, scale):
    r = z.shape[1]
    z3, d3 = z.reshape(B, T, r), dz.reshape(B, T, r)
    q, tok, ctx = pl.reshape(B, P, r), z3[:, P + 1:], z3[:, :P]
    dctx = d3[:, :P].clone()
    p = torch.exp(q @ tok.transpose(-1, -2) * scale - lse.reshape(B, P, 1))
    ds = p * (dctx @ tok.transpose(-1, -2) - (dctx * ctx).sum(-1, keepdim=True)) * scale
    d3[:, :P] = ds @ tok
    d3[:, P + 1:] += ds.transpose(-1, -2) @ q + p.transpose(-1, -2) @ dctx
    return dz


def gate_scale(x, gate):
    return gate[0] * x


def gate_grads(x, dy, gate, dx, dgate):
    dx.add_(gate[0] * dy)
    dgate.add_((x * dy).sum())


# ---------------------------------------------------------------------------------------------- head / loss
def _head(x, B, T, ps, pc, gamma, beta, wh, bh, eps, ssf_scale, ssf_shift):
    rows = x.reshape(B, T, -1)[:, ps:ps + pc]
    y = F.layer_norm(rows, (x.shape[1],), gamma, beta, eps)
    if ssf_scale is not None:
        y = y * ssf_scale + ssf_shift
    pooled = y.mean(1)
    return pooled, pooled @ wh.t() + bh


def head_fwd(x, B, T, pool_start, pool_count, gamma, beta, wh, bh, *, eps=1e-5, ssf_scale=None, ssf_shift=None):
    pooled, logits = _head(x, B, T, pool_start, pool_count, gamma, beta, wh, bh, eps, ssf_scale, ssf_shift)
    return logits, pooled


def head_bwd(x, B, T, pool_start, pool_count, gamma, beta, wh, bh, pooled, dlogits, *, dx=None, dx_lp=None, eps=1e-5, ssf_scale=None, ssf_shift=None,
             dgamma=None, dbeta=None, dssf_scale=None, dssf_shift=None, need_dx=True, dwh=None, dbh=None):
    leaves = dict(x=x.clone().requires_grad_(True), gamma=gamma.clone().requires_grad_(True), beta=beta.clone().requires_grad_(True),
                  wh=wh.clone().requires_grad_(True), bh=bh.clone().requires_grad_(True))
    if ssf_scale is not None:
        leaves.update(ss=ssf_scale.clone().requires_grad_(True), sh=ssf_shift.clone().requires_grad_(True))
    with torch.enable_grad():
        _, logits = _head(leaves['x'], B, T, pool_start, pool_count, leaves['gamma'], leaves['beta'], leaves['wh'], leaves['bh'], eps, leaves.get('ss'), leaves.get('sh'))
        gs = dict(zip(leaves, torch.autograd.grad(logits, list(leaves.values()), dlogits)))
    if need_dx:
        if dx is None:
            dx = torch.zeros_like(x)
        d3, g3 = dx.reshape(B, T, -1), gs['x'].reshape(B, T, -1)
        d3[:, pool_start:pool_start + pool_count] = g3[:, pool_start:pool_start + pool_count]
        if dx_lp is not None:
            dx_lp.reshape(B, T, -1)[:, pool_start:pool_start + pool_count] = g3[:, pool_start:pool_start + pool_count]
    if dwh is not None:
        dwh.add_(gs['wh'])
        dbh.add_(gs['bh'])
    else:
        dwh, dbh = gs['wh'], gs['bh']
    for acc, k in ((dgamma, 'gamma'), (dbeta, 'beta'), (dssf_scale, 'ss'), (dssf_shift, 'sh')):
        if acc is not None:
            acc.add_(gs[k])
    return dx, dwh, dbh


def loss_fwd_bwd(logits, target, kind, gamma=1.2, eps=1e-16, ignore_index=-100, need_grad=True):
    z = logits.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        if kind == LOSS_CE:
            loss = F.cross_entropy(z, target, ignore_index=ignore_index)
        else:
            p1 = torch.softmax(torch.clamp(z, eps, 1 - eps), -1)
            p2 = torch.softmax(torch.clamp(p1, eps, 1 - eps), -1)
            mask = target == ignore_index
            pt = p2.gather(1, (target * (~mask)).view(-1, 1)).squeeze(1) * (~mask)
            loss = ((1 - pt) ** gamma * (-torch.log(eps + pt)).masked_fill(mask, 0)).sum() / (~mask).sum()
        dz = torch.autograd.grad(loss, z)[0] if need_grad else None
    return loss.detach(), dz


# ---------------------------------------------------------------------------------------------- installation
@contextlib.contextmanager
def install(monkeypatch=None):
    """Route the engines' (and the loss module's) `ops` through this double and lift the CUDA-only guard for the duration of a test."""
    import sys
    import gaviko_b200._lib as L
    import gaviko_b200.dvpt_engine as de
    import gaviko_b200.engine as ge
    import gaviko_b200.losses.focal_loss as fl
    import gaviko_b200.vit_engine as ve
    me = sys.modules[__name__]
    saved = [(m, m.ops) for m in (ge, ve, de, fl)]
    guard = (L.require_cuda, L.device_guard)
    for m, _ in saved:
        m.ops = me
    L.require_cuda = lambda t, what='': None
    L.device_guard = lambda t: contextlib.nullcontext()
    try:
        yield
    finally:
        for m, o in saved:
            m.ops = o
        L.require_cuda, L.device_guard = guard
