"""Thin torch-tensor wrappers over the C ABI (include/gvk.h).  Plumbing only: pointer / shape marshalling and output
allocation; all arithmetic happens in libgvk_sm100a.so on torch's current CUDA stream.  No CPU fallback."""
import ctypes as C

import torch

from . import _lib as L
from ._lib import (ACT_GELU, ACT_GELU_BWD, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, ACT_NONE, LOSS_CE, LOSS_FOCAL, ROWACT_NONE, ROWACT_QUICKGELU,  # noqa: F401
                   ROWACT_RELU, GvkError)

S = L.STRUCTS

# bench.py sets this to a list to time every GEMM launch with CUDA events on the launch stream: entries (start, end, flops)
GEMM_HOOK = None

PREC_FP32, PREC_TF32 = 0, 1
# debugging aid: GVK_TF32_OFF=down,up,wgrad forces the exact-fp32 kernels for those rank-r products even in bf16 mode
import os as _os
_TF32_OFF = set(x for x in _os.environ.get('GVK_TF32_OFF', '').split(',') if x)


def _ld(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise GvkError('expected a row-major 2-D tensor (last stride 1)')
    return t.stride(0)


_Tensor, _Parameter = torch.Tensor, torch.nn.Parameter


# Device-resident replay counter mixed into every dropout seed inside the kernels (gvk.h: seed_salt).  None in eager mode; gaviko_b200.graph sets it
# to a 1-element int64 CUDA tensor so that a captured CUDA graph draws fresh masks on every replay.
SEED_SALT = None


def _set(p, **kw):
    """Fill a parameter struct; tensors become device addresses, None leaves the field zero.  (The class identity test first: isinstance() on
    torch.Tensor goes through a Python-level __instancecheck__ and was 9 of the 13 us this helper cost per call, x 564 calls per step.)"""
    for k, v in kw.items():
        if v is None:
            continue
        c = v.__class__
        if c is int or c is float:
            setattr(p, k, v)
        elif c is _Tensor or c is _Parameter or isinstance(v, _Tensor):
            setattr(p, k, L.ptr(v))
        else:
            setattr(p, k, v)
    return p


# ---------------------------------------------------------------------------------------------- GEMM
def gemm(a, b, *, out=None, out_dtype=None, bias=None, ssf_scale=None, ssf_shift=None, act=ACT_NONE, aux=None,
         pos=None, rows_per_batch=0, out_batch_rows=0, out_row_offset=0, res1=None, res2=None, out2=None, out_rows=None):
    """out[row(m), n] = epilogue(sum_k a[m,k] * b[n,k]) — see gvk_gemm in include/gvk.h."""
    M, K = a.shape
    N, Kb = b.shape
    if K != Kb or a.dtype != b.dtype:
        raise GvkError(f'gemm: operand mismatch {tuple(a.shape)} {a.dtype} x {tuple(b.shape)} {b.dtype}')
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, N), device=a.device, dtype=out_dtype or torch.float32)
    p = S['gvk_gemm_params']()
    _set(p, a=a, b=b, ab_dtype=L.dtype_tag(a.dtype), M=M, N=N, K=K, lda=_ld(a), ldb=_ld(b),
         bias=L.fptr(bias), ssf_scale=L.fptr(ssf_scale), ssf_shift=L.fptr(ssf_shift), act=act, pos=L.fptr(pos),
         rows_per_batch=rows_per_batch, out_batch_rows=out_batch_rows, out_row_offset=out_row_offset,
         out=out, out_dtype=L.dtype_tag(out.dtype), ld_out=_ld(out))
    if aux is not None:
        _set(p, aux=aux, aux_dtype=L.dtype_tag(aux.dtype), ld_aux=_ld(aux))
    if res1 is not None:
        _set(p, res1=L.ptr(res1, torch.float32), ld_res1=_ld(res1))
    if res2 is not None:
        _set(p, res2=L.ptr(res2, torch.float32), ld_res2=_ld(res2))
    if out2 is not None:
        _set(p, out2=L.ptr(out2, torch.float32), ld_out2=_ld(out2))
    if GEMM_HOOK is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.call('gvk_gemm', C.byref(p), L.stream())
        e1.record()
        GEMM_HOOK.append((e0, e1, 2.0 * M * N * K))
        return out
    L.call('gvk_gemm', C.byref(p), L.stream())
    return out


# ---------------------------------------------------------------------------------------------- row kernels
def layernorm_fwd(x, gamma, beta, *, out=None, out_dtype=torch.float32, eps=1e-5, ssf_scale=None, ssf_shift=None, save_stats=True):
    M, dim = x.shape
    if out is None:
        out = torch.empty((M, dim), device=x.device, dtype=out_dtype)
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    p = S['gvk_layernorm_fwd_params']()
    _set(p, x=L.ptr(x, torch.float32), ldx=_ld(x), gamma=L.fptr(gamma), beta=L.fptr(beta), eps=eps, ssf_scale=L.fptr(ssf_scale),
         ssf_shift=L.fptr(ssf_shift), y=out, y_dtype=L.dtype_tag(out.dtype), ldy=_ld(out), mean=mean, rstd=rstd, M=M, dim=dim)
    L.call('gvk_layernorm_fwd', C.byref(p), L.stream())
    return out, mean, rstd


def layernorm_fwd_down(x, gamma, beta, w, bias=None, *, act=ROWACT_NONE, save_pre=False, eps=1e-5, save_stats=True):
    """One pass over x (bf16 compute mode): y = LN(x) gamma + beta as bf16 (+ mean, rstd) and z = act(x @ w^T + bias) on the RAW rows
    (gvk_layernorm_fwd_down; w [r, dim]).  Returns (y, mean, rstd, dict(z, pre))."""
    M, dim = x.shape
    r = w.shape[0]
    sj, sc = _wstrides(w, r, dim, False)
    y = torch.empty((M, dim), device=x.device, dtype=torch.bfloat16)
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    z = torch.empty((M, r), device=x.device, dtype=torch.float32)
    pre = torch.empty_like(z) if save_pre else None
    p = S['gvk_layernorm_fwd_down_params']()
    _set(p, x=L.ptr(x, torch.float32), ldx=_ld(x), M=M, dim=dim, gamma=L.fptr(gamma), beta=L.fptr(beta), eps=eps, y=y, ldy=dim, mean=mean, rstd=rstd,
         w=L.ptr(w, torch.float32), w_sj=sj, w_sc=sc, bias=L.fptr(bias), r=r, act=act, pre=pre, z=z, ldz=r)
    L.call('gvk_layernorm_fwd_down', C.byref(p), L.stream())
    return y, mean, rstd, dict(z=z, pre=pre)


def layernorm_fwd_down_supported(x, r):
    return x.shape[1] in (384, 768) and r <= 32 and 'down' not in _TF32_OFF


def _wstrides(w, r, dim, transposed):
    """(w_sj, w_sc) of element (j, c) for an nn.Linear weight: [r, dim] (transposed=False) or [dim, r] (True)."""
    if not w.is_contiguous():
        raise GvkError('weights must be contiguous')
    if transposed:
        assert tuple(w.shape) == (dim, r), (tuple(w.shape), dim, r)
        return 1, r
    assert tuple(w.shape) == (r, dim), (tuple(w.shape), r, dim)
    return dim, 1


def rowproj_down(x, w, bias=None, *, transposed=False, ln=None, eps=1e-5, act=ROWACT_NONE, save_pre=False, w2=None,
                 drop_p=0.0, seed=0, offset=0, prec=PREC_FP32):
    """z = act(f(x) @ W^T + b) (W = w, or w^T when `transposed`); returns dict(z, pre, z2, mean, rstd)."""
    M, dim = x.shape
    r = w.shape[1] if transposed else w.shape[0]
    if 'down' in _TF32_OFF:
        prec = PREC_FP32
    sj, sc = _wstrides(w, r, dim, transposed)
    z = torch.empty((M, r), device=x.device, dtype=torch.float32)
    pre = torch.empty_like(z) if save_pre else None
    p = S['gvk_rowproj_down_params']()
    _set(p, x=L.ptr(x, torch.float32), ldx=_ld(x), M=M, dim=dim, r=r, w=L.ptr(w, torch.float32), w_sj=sj, w_sc=sc, bias=L.fptr(bias),
         act=act, pre=pre, z=z, ldz=r, eps=eps, drop_p=drop_p, seed=seed, offset=offset, seed_salt=SEED_SALT, precision=prec)
    mean = rstd = z2 = None
    if ln is not None:
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty(M, device=x.device, dtype=torch.float32)
        _set(p, ln_gamma=L.fptr(ln[0]), ln_beta=L.fptr(ln[1]), mean=mean, rstd=rstd)
    if w2 is not None:
        r2 = w2.shape[0]
        assert w2.shape[1] == r and w2.is_contiguous()
        z2 = torch.empty((M, r2), device=x.device, dtype=torch.float32)
        _set(p, w2=L.ptr(w2, torch.float32), r2=r2, z2=z2, ldz2=r2)
    L.call('gvk_rowproj_down', C.byref(p), L.stream())
    return dict(z=z, pre=pre, z2=z2, mean=mean, rstd=rstd)


def rowproj_up(c, w, bias=None, *, transposed=False, res=None, out=None, out_lp=None, drop_p=0.0, seed=0, offset=0, prec=PREC_FP32):
    """out = res + dropout(c @ W + b): W element (j, col) = w[col, j] for an nn.Linear(r, dim).weight (transposed=False here means
    `w` is [dim, r]); transposed=True takes a [r, dim] weight (dgrad of a down-projection)."""
    M, r = c.shape
    dim = w.shape[1] if transposed else w.shape[0]
    if 'up' in _TF32_OFF:
        prec = PREC_FP32
    if transposed:
        assert tuple(w.shape) == (r, dim) and w.is_contiguous()
        sj, sc = dim, 1
    else:
        assert tuple(w.shape) == (dim, r) and w.is_contiguous()
        sj, sc = 1, r
    if out is None:
        out = torch.empty((M, dim), device=c.device, dtype=torch.float32)
    p = S['gvk_rowproj_up_params']()
    _set(p, c=L.ptr(c, torch.float32), ldc=_ld(c), M=M, dim=dim, r=r, w=L.ptr(w, torch.float32), w_sj=sj, w_sc=sc, bias=L.fptr(bias),
         out=L.ptr(out, torch.float32), ld_out=_ld(out), drop_p=drop_p, seed=seed, offset=offset, seed_salt=SEED_SALT, precision=prec)
    if res is not None:
        _set(p, res=L.ptr(res, torch.float32), ld_res=_ld(res))
    if out_lp is not None:
        _set(p, out_lp=L.ptr(out_lp, torch.bfloat16), ld_out_lp=_ld(out_lp))
    L.call('gvk_rowproj_up', C.byref(p), L.stream())
    return out


def rowproj_up_down(c, w, bias=None, *, transposed=False, res=None, out=None, up_drop_p=0.0, up_seed=0, up_offset=0, w2=None, bias2=None, transposed2=False,
                    act=ROWACT_NONE, save_pre=False, dn_drop_p=0.0, dn_seed=0, dn_offset=0):
    """rowproj_up then rowproj_down of its output rows in one pass (gvk_rowproj_up_down; bf16 compute mode only): returns (out, dict(z, pre)).
    `w` / `transposed` as in rowproj_up, `w2` / `transposed2` as in rowproj_down."""
    M, r = c.shape
    dim = w.shape[1] if transposed else w.shape[0]
    if not (0.0 <= up_drop_p < 1.0 and 0.0 <= dn_drop_p < 1.0):
        raise GvkError(f'rowproj_up_down: dropout probabilities must be in [0, 1), got {up_drop_p}, {dn_drop_p}')
    if transposed:
        assert tuple(w.shape) == (r, dim) and w.is_contiguous()
        sj, sc = dim, 1
    else:
        assert tuple(w.shape) == (dim, r) and w.is_contiguous()
        sj, sc = 1, r
    r2 = w2.shape[1] if transposed2 else w2.shape[0]
    sj2, sc2 = _wstrides(w2, r2, dim, transposed2)
    if out is None:
        out = torch.empty((M, dim), device=c.device, dtype=torch.float32)
    z = torch.empty((M, r2), device=c.device, dtype=torch.float32)
    pre = torch.empty_like(z) if save_pre else None
    p = S['gvk_rowproj_up_down_params']()
    _set(p, c=L.ptr(c, torch.float32), ldc=_ld(c), M=M, dim=dim, r=r, w=L.ptr(w, torch.float32), w_sj=sj, w_sc=sc, bias=L.fptr(bias),
         out=L.ptr(out, torch.float32), ld_out=_ld(out), up_drop_p=up_drop_p, up_seed=up_seed, up_offset=up_offset,
         w2=L.ptr(w2, torch.float32), w2_sj=sj2, w2_sc=sc2, bias2=L.fptr(bias2), r2=r2, act=act, pre=pre, z=z, ldz=r2,
         dn_drop_p=dn_drop_p, dn_seed=dn_seed, dn_offset=dn_offset, seed_salt=SEED_SALT)
    if res is not None:
        _set(p, res=L.ptr(res, torch.float32), ld_res=_ld(res))
    L.call('gvk_rowproj_up_down', C.byref(p), L.stream())
    return out, dict(z=z, pre=pre)


def rowproj_up_down_supported(dim, r, r2, prec):
    return prec == PREC_TF32 and dim in (384, 768) and r <= 24 and r2 <= 24 and not ({'up', 'down'} & _TF32_OFF)


def skinny_wgrad(a, x, *, dw=None, dw_layout='rd', da_colsum=None, dx_colsum=None, ln=None, drop_p=0.0, seed=0, offset=0, prec=PREC_FP32, dw_strides=None):
    """dw(j,c) += sum_m a[m,j] f(x[m,c]).  dw_layout 'rd': dw is [r, dim]; 'dr': dw is [dim, r].  Accumulates (zero first)."""
    M, r = a.shape
    dim = x.shape[1]
    if 'wgrad' in _TF32_OFF:
        prec = PREC_FP32
    p = S['gvk_skinny_wgrad_params']()
    _set(p, a=L.ptr(a, torch.float32), lda=_ld(a), r=r, x=L.ptr(x, torch.float32), ldx=_ld(x), dim=dim, M=M,
         da_colsum=L.fptr(da_colsum), dx_colsum=L.fptr(dx_colsum), drop_p=drop_p, seed=seed, offset=offset, seed_salt=SEED_SALT, precision=prec)
    if dw is not None and dw_strides is not None:      # explicit (sj, sc) element strides: a column / row slice of a larger gradient
        _set(p, dw=L.ptr(dw, torch.float32), dw_sj=dw_strides[0], dw_sc=dw_strides[1])
    elif dw is not None:
        if dw_layout == 'rd':
            assert tuple(dw.shape) == (r, dim) and dw.is_contiguous()
            _set(p, dw=L.ptr(dw, torch.float32), dw_sj=dim, dw_sc=1)
        else:
            assert tuple(dw.shape) == (dim, r) and dw.is_contiguous()
            _set(p, dw=L.ptr(dw, torch.float32), dw_sj=1, dw_sc=r)
    if ln is not None:
        _set(p, ln_gamma=L.fptr(ln[0]), ln_beta=L.fptr(ln[1]), mean=L.fptr(ln[2]), rstd=L.fptr(ln[3]))
    fn = L.lib().gvk_skinny_wgrad_ws_floats
    fn.restype = C.c_size_t
    n_ws = int(fn(r, dim, M))
    ws = torch.empty(n_ws, device=a.device, dtype=torch.float32)
    _set(p, ws=ws, ws_floats=n_ws)
    L.call('gvk_skinny_wgrad', C.byref(p), L.stream())


def layernorm_bwd(x, gamma, mean, rstd, *, dy=None, dz=None, w=None, dres=None, dx=None, dx_lp=None, dgamma=None, dbeta=None, az=None, aw=None,
                  beta=None, ssf_scale=None, dssf_scale=None, dssf_shift=None, prec=PREC_FP32, ow=None, ow_transposed=False):
    """dx = dres + LN'(dy) + az @ aw;  dy dense [M, dim] or rank-r (dz [M, r], w [r, dim]);  az [M, ra], aw [ra, dim].  prec = PREC_TF32: the
    rank-r product of the two forms gvk.h names runs on the tensor cores (tf32 operands) inside the pass.  ow ([r, dim], or [dim, r] with
    ow_transposed): also returns oz = dx @ W^T computed from the output rows in the same pass (tensor-core form only) -> (dx, oz)."""
    M, dim = x.shape
    if dx is None:
        dx = torch.empty((M, dim), device=x.device, dtype=torch.float32)
    p = S['gvk_layernorm_bwd_params']()
    _set(p, x=L.ptr(x, torch.float32), ldx=_ld(x), gamma=L.fptr(gamma), mean=L.fptr(mean), rstd=L.fptr(rstd),
         dx=L.ptr(dx, torch.float32), ld_dx=_ld(dx), dgamma=L.fptr(dgamma), dbeta=L.fptr(dbeta), M=M, dim=dim, precision=prec)
    if dy is not None:
        _set(p, dy=L.ptr(dy), ld_dy=_ld(dy), dy_dtype=L.dtype_tag(dy.dtype))
    if dz is not None:
        r = dz.shape[1]
        assert tuple(w.shape) == (r, dim) and w.is_contiguous()
        _set(p, dz=L.ptr(dz, torch.float32), ld_dz=_ld(dz), w=L.ptr(w, torch.float32), w_sj=dim, w_sc=1, r=r)
    if dres is not None:
        _set(p, dres=L.ptr(dres, torch.float32), ld_dres=_ld(dres))
    if dx_lp is not None:
        _set(p, dx_lp=L.ptr(dx_lp, torch.bfloat16), ld_dx_lp=_ld(dx_lp))
    if ssf_scale is not None:
        _set(p, beta=L.fptr(beta), ssf_scale=L.fptr(ssf_scale), dssf_scale=L.fptr(dssf_scale), dssf_shift=L.fptr(dssf_shift))
    if az is not None:
        ra = az.shape[1]
        assert tuple(aw.shape) == (ra, dim) and aw.is_contiguous()
        _set(p, az=L.ptr(az, torch.float32), ld_az=_ld(az), aw=L.ptr(aw, torch.float32), aw_sj=dim, aw_sc=1, ra=ra)
    oz = None
    if ow is not None:
        orank = ow.shape[1] if ow_transposed else ow.shape[0]
        sj, sc = _wstrides(ow, orank, dim, ow_transposed)
        oz = torch.empty((M, orank), device=x.device, dtype=torch.float32)
        _set(p, ow=L.ptr(ow, torch.float32), ow_sj=sj, ow_sc=sc, orank=orank, oz=oz, ld_oz=orank)
    L.call('gvk_layernorm_bwd', C.byref(p), L.stream())
    return dx if ow is None else (dx, oz)


def layernorm_bwd_down_supported(x, orank, prec):
    return prec == PREC_TF32 and x.shape[1] in (384, 768) and orank <= 24 and 'down' not in _TF32_OFF


def small_wgrad(a, b, dw):
    """dw[j, k] += sum_m a[m, j] b[m, k]"""
    assert tuple(dw.shape) == (a.shape[1], b.shape[1]) and dw.is_contiguous()
    L.call('gvk_small_wgrad', C.c_void_p(L.ptr(a, torch.float32)), _ld(a), a.shape[1], C.c_void_p(L.ptr(b, torch.float32)), _ld(b), b.shape[1],
           a.shape[0], C.c_void_p(L.fptr(dw)), L.stream())


def small_matmul(a, w):
    """a [M, ra] @ w [ra, rb] -> [M, rb]"""
    assert w.is_contiguous() and w.shape[0] == a.shape[1]
    out = torch.empty((a.shape[0], w.shape[1]), device=a.device, dtype=torch.float32)
    L.call('gvk_small_matmul', C.c_void_p(L.ptr(a, torch.float32)), _ld(a), a.shape[1], C.c_void_p(L.fptr(w)), w.shape[1], a.shape[0],
           C.c_void_p(L.fptr(out)), _ld(out), L.stream())
    return out


def colsum(x, out):
    """out[c] += sum_m x[m, c]"""
    L.call('gvk_colsum', C.c_void_p(L.ptr(x, torch.float32)), _ld(x), x.shape[0], x.shape[1], C.c_void_p(L.fptr(out)), L.stream())


def cast_bf16(x, out=None):
    M, dim = x.shape
    if out is None:
        out = torch.empty((M, dim), device=x.device, dtype=torch.bfloat16)
    L.call('gvk_cast_f32_bf16', C.c_void_p(L.ptr(x, torch.float32)), _ld(x), C.c_void_p(L.ptr(out, torch.bfloat16)), _ld(out), M, dim, L.stream())
    return out


def cast_f32(x, out=None):
    """bf16 [M, dim] (row-major view) -> fp32"""
    M, dim = x.shape
    if out is None:
        out = torch.empty((M, dim), device=x.device, dtype=torch.float32)
    L.call('gvk_cast_bf16_f32', C.c_void_p(L.ptr(x, torch.bfloat16)), _ld(x), C.c_void_p(L.ptr(out, torch.float32)), _ld(out), M, dim, L.stream())
    return out


def ssf_bwd(dy, *, y=None, scale=None, shift=None, dx=None, dscale=None, dshift=None, sub=None, rows_per_batch=0, batch_rows=0, M=None):
    """SSF site backward (scale given) or bias gradient (scale None): see gvk_ssf_bwd in include/gvk.h.  dx may be dy (in place)."""
    N = dy.shape[1]
    p = S['gvk_ssf_bwd_params']()
    _set(p, dy=dy, ld_dy=_ld(dy), dtype=L.dtype_tag(dy.dtype), M=dy.shape[0] if M is None else M, N=N, rows_per_batch=rows_per_batch, batch_rows=batch_rows,
         dscale=L.fptr(dscale), dshift=L.fptr(dshift))
    if scale is not None:
        assert y is not None and y.dtype == dy.dtype
        _set(p, y=y, ld_y=_ld(y), scale=L.fptr(scale), shift=L.fptr(shift))
        if dx is not None:
            assert dx.dtype == dy.dtype
            _set(p, dx=dx, ld_dx=_ld(dx))
    if sub is not None:
        _set(p, sub=L.ptr(sub, torch.float32), ld_sub=_ld(sub))
    L.call('gvk_ssf_bwd', C.byref(p), L.stream())
    return dx


def dropout(x, drop_p, seed, *, res=None, out=None, out_dtype=None, offset=0):
    """out = res + dropout(x) with the replayable Philox mask of element (m, n) = philox(seed, offset + m*N + n)."""
    M, N = x.shape
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=out_dtype or x.dtype)
    p = S['gvk_dropout_params']()
    _set(p, x=x, x_dtype=L.dtype_tag(x.dtype), ldx=_ld(x), out=out, out_dtype=L.dtype_tag(out.dtype), ld_out=_ld(out), M=M, N=N, drop_p=drop_p, seed=seed, offset=offset, seed_salt=SEED_SALT)
    if res is not None:
        _set(p, res=L.ptr(res, torch.float32), ld_res=_ld(res))
    L.call('gvk_dropout', C.byref(p), L.stream())
    return out


# ---------------------------------------------------------------------------------------------- attention (SIMT)
def _attn_params(qkv, B, T, H, D, q_off, k_off, v_off, scale, window, grid, drop_p, seed, offset, out, lse, prec=PREC_FP32):
    f = S['gvk_attn_fwd_params']()
    _set(f, precision=PREC_FP32 if 'attn' in _TF32_OFF else prec, qkv=qkv, dtype=L.dtype_tag(qkv.dtype), ld=_ld(qkv), q_off=q_off, k_off=k_off, v_off=v_off, B=B, T=T, H=H, D=D, scale=scale,
         drop_p=drop_p, seed=seed, offset=offset, seed_salt=SEED_SALT, out=out, ld_out=_ld(out), lse=L.fptr(lse))
    if window is not None:
        _set(f, win_d=window[0], win_h=window[1], win_w=window[2], grid_d=grid[0], grid_h=grid[1], grid_w=grid[2])
    return f


def attn_simt_fwd(qkv, B, T, H, D, *, q_off, k_off, v_off, scale, window=None, grid=None, drop_p=0.0, seed=0, offset=0, prec=PREC_FP32):
    out = torch.empty((B * T, H * D), device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty(B * H * T, device=qkv.device, dtype=torch.float32)
    f = _attn_params(qkv, B, T, H, D, q_off, k_off, v_off, scale, window, grid, drop_p, seed, offset, out, lse, prec)
    L.call('gvk_attn_simt_fwd', C.byref(f), L.stream())
    return out, lse


def attn_simt_bwd(qkv, out, lse, dout, B, T, H, D, *, q_off, k_off, v_off, scale, window=None, grid=None, drop_p=0.0, seed=0, offset=0, dqkv=None,
                  prec=PREC_FP32):
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    p = S['gvk_attn_bwd_params']()
    p.f = _attn_params(qkv, B, T, H, D, q_off, k_off, v_off, scale, window, grid, drop_p, seed, offset, out, lse, prec)
    assert dout.dtype == qkv.dtype and dqkv.dtype == qkv.dtype
    _set(p, dout=dout, ld_dout=_ld(dout), delta=delta, dqkv=dqkv, ld_dqkv=_ld(dqkv))
    L.call('gvk_attn_simt_bwd', C.byref(p), L.stream())
    return dqkv


def mhsa_fwd(qkv, B, T, H, scale, drop_p=0.0, seed=0):
    """tcgen05 flash attention forward: qkv [B*T, 3*H*64] bf16 -> (out [B*T, H*64] bf16, lse [B*H*T] fp32); optional Philox dropout on the probabilities."""
    if qkv.dtype != torch.bfloat16:
        raise GvkError('mhsa_fwd: bf16 only (fp32 mode uses attn_simt_fwd)')
    out = torch.empty((B * T, H * 64), device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(B * H * T, device=qkv.device, dtype=torch.float32)
    p = S['gvk_mhsa_fwd_params']()
    _set(p, qkv=qkv, ld=_ld(qkv), B=B, T=T, H=H, scale=scale, out=out, ld_out=_ld(out), lse=lse, drop_p=float(drop_p), seed=int(seed) & 0xFFFFFFFFFFFFFFFF, seed_salt=SEED_SALT)
    L.call('gvk_mhsa_fwd', C.byref(p), L.stream())
    return out, lse


def mhsa_bwd(qkv, out, lse, dout, B, T, H, scale, drop_p=0.0, seed=0):
    if qkv.dtype != torch.bfloat16 or dout.dtype != torch.bfloat16:
        raise GvkError('mhsa_bwd: bf16 only')
    dqkv = torch.empty_like(qkv)
    fn = L.lib().gvk_mhsa_bwd_ws_floats
    fn.restype = C.c_size_t
    delta = torch.empty(int(fn(B, T, H)), device=lse.device, dtype=torch.float32)
    p = S['gvk_mhsa_bwd_params']()
    mask = None
    if drop_p > 0.0:
        fm = L.lib().gvk_mhsa_bwd_mask_words
        fm.restype = C.c_size_t
        mask = torch.empty(int(fm(B, T, H)), device=lse.device, dtype=torch.int32)
    _set(p, qkv=qkv, ld=_ld(qkv), B=B, T=T, H=H, scale=scale, out=out, ld_out=_ld(out), lse=lse, dout=dout, ld_dout=_ld(dout), delta=delta,
         dqkv=dqkv, ld_dqkv=_ld(dqkv), drop_p=float(drop_p), seed=int(seed) & 0xFFFFFFFFFFFFFFFF, seed_salt=SEED_SALT, mask_ws=mask)
    L.call('gvk_mhsa_bwd', C.byref(p), L.stream())
    return dqkv


# ---------------------------------------------------------------------------------------------- token assembly
def rescale_intensity(x, out_min=0.0, out_max=1.0, *, out=None, out_dtype=None):
    """Per-volume (x - min) / (max - min) * (out_max - out_min) + out_min over x[b] (see gvk_rescale_intensity); x fp32 contiguous, leading
    dimension = volumes.  out may be x itself (fp32)."""
    if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() < 2:
        raise GvkError('rescale_intensity: expected a contiguous fp32 tensor (B, ...)')
    B = x.shape[0]
    n = x.numel() // B
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=out_dtype or torch.float32)
    ws = torch.empty(2 * 64 * B, device=x.device, dtype=torch.float32)
    p = S['gvk_rescale_intensity_params']()
    _set(p, out=out, out_dtype=L.dtype_tag(out.dtype), B=B, n=n, out_min=out_min, out_max=out_max, ws=ws)
    setattr(p, 'in', L.ptr(x, torch.float32))       # 'in' is a Python keyword
    L.call('gvk_rescale_intensity', C.byref(p), L.stream())
    return out


def split_pack_bf16(src, dst, pattern):
    """dst [rows, width] bf16 (a view with a row stride is fine) <- hi / lo split of src [rows, r] fp32 in three r-wide slots (see gvk_split_pack_bf16)."""
    rows, r = src.shape
    if dst.dtype != torch.bfloat16 or dst.shape[0] != rows or dst.shape[1] < 3 * r:
        raise GvkError('split_pack_bf16: dst must be bf16 [rows, >= 3 r]')
    L.call('gvk_split_pack_bf16', C.c_void_p(L.ptr(src, torch.float32)), _ld(src), rows, r, C.c_void_p(L.ptr(dst)), _ld(dst), dst.shape[1], pattern, L.stream())
    return dst


def patch_gather(img, fp, ps, out_dtype):
    B, Cc, D, H, W = img.shape
    if img.dtype != torch.float32 or not img.is_contiguous():
        raise GvkError('patch_gather: expected a contiguous fp32 (B, C, D, H, W) volume')
    n = (D // fp) * (H // ps) * (W // ps)
    out = torch.empty((B * n, Cc * fp * ps * ps), device=img.device, dtype=out_dtype)
    L.call('gvk_patch_gather', C.c_void_p(L.ptr(img)), B, Cc, D, H, W, fp, ps, C.c_void_p(L.ptr(out)), L.dtype_tag(out_dtype), L.stream())
    return out


def patch_embed(img, fp, ps, weight, bias, pos, out, out_batch_rows, out_row_offset, out2=None):
    """Fused TMA patch gather + tf32 tcgen05 GEMM + bias / positional embedding / token-row scatter (gvk_patch_embed).  Returns False (and does
    nothing) when the geometry is outside the kernel's range, so that the caller can take the gather + GEMM route."""
    B, Cc, D, H, W = img.shape
    if img.dtype != torch.float32 or not img.is_contiguous() or weight.dtype != torch.float32 or not weight.is_contiguous():
        raise GvkError('patch_embed: expected contiguous fp32 volume and weight')
    p = S['gvk_patch_embed_params']()
    _set(p, img=img, B=B, C=Cc, D=D, H=H, W=W, fp=fp, ps=ps, weight=weight, bias=bias, pos=pos, dim=weight.shape[0], out=out, ld_out=_ld(out),
         out_batch_rows=out_batch_rows, out_row_offset=out_row_offset, out2=out2, ld_out2=_ld(out2) if out2 is not None else 0)
    if not L.lib().gvk_patch_embed_supported(C.byref(p)):
        return False
    L.call('gvk_patch_embed', C.byref(p), L.stream())
    return True


# ---------------------------------------------------------------------------------------------- EVP primitives (csrc/gvk_evp.cu)
def wgrad(a, b, dw, *, M=None, a_rows=None, b_rows=None, prec=PREC_FP32):
    """dw[i, j] += sum_m a[row_a(m), i] * b[row_b(m), j]  (gvk_wgrad).  a_rows / b_rows = (rows_per_batch, batch_rows) row maps (None:
    identity); M = logical row count (default: a.shape[0]).  a, b: fp32 or bf16 row-major; dw: fp32 row-major view (accumulated into)."""
    if dw.dtype != torch.float32 or tuple(dw.shape) != (a.shape[1], b.shape[1]):
        raise GvkError(f'wgrad: dw must be fp32 {(a.shape[1], b.shape[1])}, got {dw.dtype} {tuple(dw.shape)}')
    ar, br = a_rows or (0, 0), b_rows or (0, 0)
    p = S['gvk_wgrad_params']()
    _set(p, a=a, a_dtype=L.dtype_tag(a.dtype), lda=_ld(a), na=a.shape[1], a_rows_per_batch=ar[0], a_batch_rows=ar[1],
         b=b, b_dtype=L.dtype_tag(b.dtype), ldb=_ld(b), nb=b.shape[1], b_rows_per_batch=br[0], b_batch_rows=br[1],
         M=a.shape[0] if M is None else M, dw=dw, ld_dw=_ld(dw), precision=PREC_FP32 if 'wgrad' in _TF32_OFF else prec)
    L.call('gvk_wgrad', C.byref(p), L.stream())
    return dw


def hfreq_filter(img, filt, hit):
    """| filt @ slice | on the depth slices with hit[d] != 0, | slice | elsewhere (gvk_hfreq_filter; reference model/evp.py:124-146)."""
    B, Cc, D, H, W = img.shape
    if img.dtype != torch.float32 or not img.is_contiguous():
        raise GvkError('hfreq_filter: expected a contiguous fp32 (B, C, D, H, W) volume')
    if tuple(filt.shape) != (H, H) or not filt.is_contiguous() or hit.dtype != torch.uint8 or hit.numel() != D:
        raise GvkError('hfreq_filter: filt must be a contiguous fp32 [H, H] matrix and hit a uint8 [D] vector')
    out = torch.empty_like(img)
    p = S['gvk_hfreq_filter_params']()
    n = B * Cc * D
    step = 65535 // D * D      # slices per launch (grid.z limit), a whole number of depth stacks
    xi, xo = img.view(n, H, W), out.view(n, H, W)
    for s0 in range(0, n, step):
        s1 = min(n, s0 + step)
        _set(p, out=xo[s0:s1], filt=L.fptr(filt), hit=L.ptr(hit), slices=s1 - s0, D=D, H=H, W=W)
        setattr(p, 'in', L.ptr(xi[s0:s1]))
        L.call('gvk_hfreq_filter', C.byref(p), L.stream())
    return out


def fill_rows(a, b, out, out_batch_rows, out_row_offset, B):
    R, dim = a.shape
    L.call('gvk_fill_rows', C.c_void_p(L.fptr(a)), C.c_void_p(L.fptr(b)) if b is not None else None, R, dim, C.c_void_p(L.ptr(out, torch.float32)), _ld(out),
           out_batch_rows, out_row_offset, B, L.stream())


def batch_rowsum(x, batch_rows, row_offset, R, B, out=None, accumulate=False):
    dim = x.shape[1]
    if out is None:
        out = torch.empty((R, dim), device=x.device, dtype=torch.float32)
    L.call('gvk_batch_rowsum', C.c_void_p(L.ptr(x, torch.float32)), _ld(x), batch_rows, row_offset, R, dim, B, C.c_void_p(L.fptr(out)), int(accumulate), L.stream())
    return out


# ---------------------------------------------------------------------------------------------- prompt fusion
FUSION_WEIGHT_FIELDS = ('wq_g', 'bq_g', 'wq_l', 'bq_l', 'a_ln_w', 'a_ln_b', 'a_w1', 'a_b1', 'a_w3', 'a_b3', 'g_ln_w', 'g_ln_b', 'g_w', 'g_b')
FUSION_SAVED_FIELDS = ('pl', 'qg', 'ql', 'ctx_g', 'ctx_l', 'lse_g', 'lse_l', 'imp', 'gw')


def _fusion_weights(w):
    s = S['gvk_fusion_weights']()
    for k in FUSION_WEIGHT_FIELDS:
        setattr(s, k, L.fptr(w[k]))
    return s


def prompt_fusion_fwd(xl, ll, w, B, T, N, P):
    """xl [B*T, r] is modified in place (rows < P of every volume become the enhanced prompts).  Returns the saved state."""
    r = xl.shape[1]
    dev = xl.device
    saved = {k: torch.empty((B * P, r), device=dev, dtype=torch.float32) for k in ('pl', 'qg', 'ql', 'ctx_g', 'ctx_l')}
    saved.update({k: torch.empty(B * P, device=dev, dtype=torch.float32) for k in ('lse_g', 'lse_l', 'imp')})
    saved['gw'] = torch.empty(B, device=dev, dtype=torch.float32)
    p = S['gvk_fusion_fwd_params']()
    _set(p, xl=L.fptr(xl), ll=L.fptr(ll), B=B, T=T, N=N, P=P, r=r)
    p.w = _fusion_weights(w)
    for k in FUSION_SAVED_FIELDS:
        setattr(p.s, k, L.fptr(saved[k]))
    L.call('gvk_prompt_fusion_fwd', C.byref(p), L.stream())
    return saved


def prompt_fusion_bwd(xl, ll, dxl, w, saved, grads, B, T, N, P):
    """dxl [B*T, r]: d(combined) in, d(xl) out (in place).  Returns dll [B*N, r].  `grads` (dict of zeroed tensors) accumulates."""
    r = xl.shape[1]
    dll = torch.empty((B * N, r), device=xl.device, dtype=torch.float32)
    ws = torch.empty(B * P * (2 * r + 4), device=xl.device, dtype=torch.float32)
    p = S['gvk_fusion_bwd_params']()
    _set(p, xl=L.fptr(xl), ll=L.fptr(ll), dxl=L.fptr(dxl), dll=L.fptr(dll), B=B, T=T, N=N, P=P, r=r, ws=L.fptr(ws))
    p.w = _fusion_weights(w)
    for k in FUSION_SAVED_FIELDS:
        setattr(p.s, k, L.fptr(saved[k]))
    for k in FUSION_WEIGHT_FIELDS:
        setattr(p.g, k, L.fptr(grads[k]))
    L.call('gvk_prompt_fusion_bwd', C.byref(p), L.stream())
    return dll


def relu_bwd(dy, z, out=None):
    if out is None:
        out = torch.empty_like(dy)
    L.call('gvk_relu_bwd', C.c_void_p(L.fptr(dy)), C.c_void_p(L.fptr(z)), C.c_void_p(L.fptr(out)), C.c_size_t(dy.numel()), L.stream())
    return out


def quickgelu_bwd(dy, pre, out=None):
    if out is None:
        out = torch.empty_like(dy)
    L.call('gvk_quickgelu_bwd', C.c_void_p(L.fptr(dy)), C.c_void_p(L.fptr(pre)), C.c_void_p(L.fptr(out)), C.c_size_t(dy.numel()), L.stream())
    return out


# ---------------------------------------------------------------------------------------------- DVPT side path
def quickgelu_fwd(x, out=None):
    if out is None:
        out = torch.empty_like(x)
    L.call('gvk_quickgelu_fwd', C.c_void_p(L.fptr(x)), C.c_void_p(L.fptr(out)), C.c_size_t(x.numel()), L.stream())
    return out


def quickgelu_bwd_add(dy, pre, res=None, out=None):
    """out = res + dy * quickgelu'(pre)"""
    if out is None:
        out = torch.empty_like(dy)
    L.call('gvk_quickgelu_bwd_add', C.c_void_p(L.fptr(dy)), C.c_void_p(L.fptr(pre)), C.c_void_p(L.fptr(res)) if res is not None else None,
           C.c_void_p(L.fptr(out)), C.c_size_t(dy.numel()), L.stream())
    return out


def latent_xattn_fwd(z, B, T, P, scale):
    """z [B*T, r] modified in place (prompt rows <- attention output); returns (pl [B*P, r], lse [B*P])."""
    r = z.shape[1]
    pl = torch.empty((B * P, r), device=z.device, dtype=torch.float32)
    lse = torch.empty(B * P, device=z.device, dtype=torch.float32)
    p = S['gvk_latent_xattn_fwd_params']()
    _set(p, z=L.fptr(z), B=B, T=T, P=P, r=r, scale=scale, pl=pl, lse=lse)
    L.call('gvk_latent_xattn_fwd', C.byref(p), L.stream())
    return pl, lse


def latent_xattn_bwd(z, pl, lse, dz, B, T, P, scale):
    """dz [B*T, r]: d(combined) in, d(latent) out (in place)."""
    p = S['gvk_latent_xattn_bwd_params']()
    _set(p, z=L.fptr(z), pl=L.fptr(pl), lse=L.fptr(lse), dz=L.fptr(dz), B=B, T=T, P=P, r=z.shape[1], scale=scale)
    L.call('gvk_latent_xattn_bwd', C.byref(p), L.stream())
    return dz


def gate_scale(x, gate):
    out = torch.empty_like(x)
    L.call('gvk_gate_scale', C.c_void_p(L.fptr(x)), C.c_void_p(L.fptr(gate)), C.c_void_p(L.fptr(out)), C.c_size_t(x.numel()), L.stream())
    return out


def gate_grads(x, dy, gate, dx, dgate):
    """dx += gate * dy; dgate += <x, dy>"""
    L.call('gvk_gate_grads', C.c_void_p(L.fptr(x)), C.c_void_p(L.fptr(dy)), C.c_void_p(L.fptr(gate)), C.c_void_p(L.fptr(dx)), C.c_void_p(L.fptr(dgate)),
           C.c_size_t(x.numel()), L.stream())


# ---------------------------------------------------------------------------------------------- head / loss
def _head_params(x, B, T, pool_start, pool_count, gamma, beta, wh, bh, pooled, logits, eps, ssf_scale, ssf_shift):
    f = S['gvk_head_fwd_params']()
    _set(f, x=L.ptr(x, torch.float32), ldx=_ld(x), B=B, T=T, dim=x.shape[1], pool_start=pool_start, pool_count=pool_count, gamma=L.fptr(gamma),
         beta=L.fptr(beta), eps=eps, ssf_scale=L.fptr(ssf_scale), ssf_shift=L.fptr(ssf_shift), wh=L.fptr(wh), bh=L.fptr(bh),
         num_classes=wh.shape[0], pooled=L.fptr(pooled), logits=L.fptr(logits))
    return f


def head_fwd(x, B, T, pool_start, pool_count, gamma, beta, wh, bh, *, eps=1e-5, ssf_scale=None, ssf_shift=None):
    pooled = torch.empty((B, x.shape[1]), device=x.device, dtype=torch.float32)
    logits = torch.empty((B, wh.shape[0]), device=x.device, dtype=torch.float32)
    f = _head_params(x, B, T, pool_start, pool_count, gamma, beta, wh, bh, pooled, logits, eps, ssf_scale, ssf_shift)
    L.call('gvk_head_fwd', C.byref(f), L.stream())
    return logits, pooled


def head_bwd(x, B, T, pool_start, pool_count, gamma, beta, wh, bh, pooled, dlogits, *, dx=None, dx_lp=None, eps=1e-5, ssf_scale=None,
             ssf_shift=None, dgamma=None, dbeta=None, dssf_scale=None, dssf_shift=None, need_dx=True, dwh=None, dbh=None):
    """Returns (dx, dwh, dbh).  dx must be pre-zeroed when pool_count < T (only pooled rows are written)."""
    dim = x.shape[1]
    if need_dx and dx is None:
        dx = torch.zeros((B * T, dim), device=x.device, dtype=torch.float32)
    accumulate = dwh is not None
    if not accumulate:
        dwh = torch.empty_like(wh)
        dbh = torch.empty_like(bh)
    p = S['gvk_head_bwd_params']()
    p.accumulate_w = int(accumulate)
    p.f = _head_params(x, B, T, pool_start, pool_count, gamma, beta, wh, bh, pooled, None, eps, ssf_scale, ssf_shift)
    _set(p, dlogits=L.fptr(dlogits), dwh=L.fptr(dwh), dbh=L.fptr(dbh), dgamma=L.fptr(dgamma), dbeta=L.fptr(dbeta),
         dssf_scale=L.fptr(dssf_scale), dssf_shift=L.fptr(dssf_shift))
    if need_dx:
        _set(p, dx=L.ptr(dx, torch.float32), ld_dx=_ld(dx))
        if dx_lp is not None:
            _set(p, dx_lp=L.ptr(dx_lp, torch.bfloat16), ld_dx_lp=_ld(dx_lp))
    L.call('gvk_head_bwd', C.byref(p), L.stream())
    return dx, dwh, dbh


def loss_fwd_bwd(logits, target, kind, gamma=1.2, eps=1e-16, ignore_index=-100, need_grad=True):
    B, Cn = logits.shape
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits) if need_grad else None
    if target.dtype != torch.int64 or not target.is_contiguous():
        raise GvkError('loss: target must be a contiguous int64 tensor')
    L.call('gvk_loss_fwd_bwd', C.c_void_p(L.fptr(logits)), C.c_void_p(L.ptr(target)), B, Cn, kind, C.c_float(gamma), C.c_float(eps),
           C.c_longlong(ignore_index), C.c_void_p(L.fptr(loss)), C.c_void_p(L.fptr(dlogits)) if need_grad else None, L.stream())
    return loss, dlogits
