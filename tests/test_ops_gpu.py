"""GPU: every side-path kernel against a plain torch fp32/fp64 restatement (oracle functions where they exist)."""
import pytest
import torch
import torch.nn.functional as F

from gaviko_b200 import ops
from oracle import gaviko_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def close(a, b, tol=2e-5):
    a, b = a.double(), b.double()
    scale = max(1.0, b.abs().max().item())
    err = (a - b).abs().max().item()
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize('dim', [192, 768, 1024])
def test_layernorm_fwd_bwd(dim):
    torch.manual_seed(dim)
    M = 517
    x = torch.randn(M, dim, device=DEV) * 2 + 0.3
    g = torch.randn(dim, device=DEV) * 0.1 + 1
    b = torch.randn(dim, device=DEV) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.layer_norm(xr, (dim,), gr, br, 1e-5)
    close(y, yr.detach())
    ybf, _, _ = ops.layernorm_fwd(x, g, b, out_dtype=torch.bfloat16)
    close(ybf, yr.detach(), 1e-2)
    dy = torch.randn(M, dim, device=DEV)
    dres = torch.randn(M, dim, device=DEV)
    yr.backward(dy.double())
    dgam, dbet = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    dx_lp = torch.empty(M, dim, device=DEV, dtype=torch.bfloat16)
    dx = ops.layernorm_bwd(x, g, mean, rstd, dy=dy, dres=dres, dgamma=dgam, dbeta=dbet, dx_lp=dx_lp)
    close(dx, xr.grad + dres.double())
    close(dx_lp, xr.grad + dres.double(), 1e-2)
    close(dgam, gr.grad, 1e-4)
    close(dbet, br.grad, 1e-4)


@pytest.mark.parametrize('prec', [ops.PREC_FP32, ops.PREC_TF32])
@pytest.mark.parametrize('dim,r', [(192, 20), (768, 20), (1024, 32), (768, 4), (384, 20)])
def test_rowproj_down_up_wgrad(dim, r, prec):
    """prec = TF32: operands rounded to 10 mantissa bits (rel 4.9e-4 each), fp32 accumulation -> tolerance 3e-3 of the output scale."""
    tc = prec == ops.PREC_TF32
    t5, t4 = (3e-3, 3e-3) if tc else (2e-5, 1e-4)
    torch.manual_seed(dim + r)
    M = 1033
    x = torch.randn(M, dim, device=DEV)
    w = torch.randn(r, dim, device=DEV) / dim ** 0.5
    b = torch.randn(r, device=DEV) * 0.1
    g = torch.randn(dim, device=DEV) * 0.1 + 1
    be = torch.randn(dim, device=DEV) * 0.1
    w2 = torch.randn(3 * r, r, device=DEV)
    # LN + down + chained projection
    d = ops.rowproj_down(x, w, b, ln=(g, be), w2=w2, prec=prec)
    zr = F.layer_norm(x.double(), (dim,), g.double(), be.double(), 1e-5) @ w.double().t() + b.double()
    close(d['z'], zr, t5)
    close(d['z2'], zr @ w2.double().t(), t4)
    # QuickGELU variant with saved pre-activation
    d2 = ops.rowproj_down(x, w, b, act=ops.ROWACT_QUICKGELU, save_pre=True, prec=prec)
    pre = x.double() @ w.double().t() + b.double()
    close(d2['pre'], pre, t5)
    close(d2['z'], pre * torch.sigmoid(1.702 * pre), t5)
    # transposed weight ([dim, r] used as dgrad of an up-projection)
    wu = torch.randn(dim, r, device=DEV) / r ** 0.5
    close(ops.rowproj_down(x, wu, transposed=True, prec=prec)['z'], x.double() @ wu.double(), t5)
    # up projection + residual (+ bf16 copy)
    c = torch.randn(M, r, device=DEV)
    bu = torch.randn(dim, device=DEV) * 0.1
    res = torch.randn(M, dim, device=DEV)
    lp = torch.empty(M, dim, device=DEV, dtype=torch.bfloat16)
    out = ops.rowproj_up(c, wu, bu, res=res, out_lp=lp, prec=prec)
    ref = c.double() @ wu.double().t() + bu.double() + res.double()
    close(out, ref, t5)
    close(lp, ref, 1e-2)
    close(ops.rowproj_up(c, w, transposed=True, prec=prec), c.double() @ w.double(), t5)
    # in-place accumulate
    acc = res.clone()
    ops.rowproj_up(c, w, transposed=True, res=acc, out=acc, prec=prec)
    close(acc, res.double() + c.double() @ w.double(), t5)
    # skinny wgrad, both layouts, with colsums, with LN recompute
    a = torch.randn(M, r, device=DEV)
    dw = torch.zeros(r, dim, device=DEV)
    dac = torch.zeros(r, device=DEV)
    dxc = torch.zeros(dim, device=DEV)
    ops.skinny_wgrad(a, x, dw=dw, da_colsum=dac, dx_colsum=dxc, prec=prec)
    close(dw, a.double().t() @ x.double(), t4)
    close(dac, a.double().sum(0), 1e-4)
    close(dxc, x.double().sum(0), t5 if tc else 1e-4)      # tf32 form: the column sum rides on the MMA (ones slot), operands rounded to tf32
    dwt = torch.zeros(dim, r, device=DEV)
    ops.skinny_wgrad(a, x, dw=dwt, dw_layout='dr', prec=prec)
    close(dwt, x.double().t() @ a.double(), t4)
    dwl = torch.zeros(r, dim, device=DEV)
    ops.skinny_wgrad(a, x, dw=dwl, ln=(g, be, d['mean'], d['rstd']), prec=prec)
    close(dwl, a.double().t() @ F.layer_norm(x.double(), (dim,), g.double(), be.double(), 1e-5), t4)
    # LN backward in rank-r form
    dz = torch.randn(M, r, device=DEV)
    xr = x.double().requires_grad_(True)
    (F.layer_norm(xr, (dim,), g.double(), be.double(), 1e-5) @ w.double().t()).backward(dz.double())
    close(ops.layernorm_bwd(x, g, d['mean'], d['rstd'], dz=dz, w=w), xr.grad, 1e-4)
    # dense dy + additive rank-r term outside the norm (+ residual), in place over dy
    dy = torch.randn(M, dim, device=DEV)
    dres = torch.randn(M, dim, device=DEV)
    xr2 = x.double().requires_grad_(True)
    F.layer_norm(xr2, (dim,), g.double(), be.double(), 1e-5).backward(dy.double())
    want = xr2.grad + dres.double() + dz.double() @ w.double()
    got = ops.layernorm_bwd(x, g, d['mean'], d['rstd'], dy=dy, dres=dres, dx=dy, az=dz, aw=w)
    close(got, want, 1e-4)


@pytest.mark.parametrize('M,dim,r', [(1033, 768, 20), (16 * 300 + 5, 768, 20), (517, 384, 20), (2066, 768, 32), (9, 768, 8), (16 * 148 * 3, 768, 20)])
def test_layernorm_bwd_tensor_core_forms(M, dim, r):
    """prec = TF32: the two rank-r forms of the GAViKO backward (gvk.h) with the rank-r product as tf32 MMAs inside the LayerNorm pass.
    Ragged row counts, more steps per CTA than the cp.async slots hold, in-place residual, parameter gradients."""
    torch.manual_seed(M + dim + r)
    x = torch.randn(M, dim, device=DEV) * 2 + 0.3
    g = torch.randn(dim, device=DEV) * 0.1 + 1
    be = torch.randn(dim, device=DEV) * 0.1
    _, mean, rstd = ops.layernorm_fwd(x, g, be)
    w = torch.randn(r, dim, device=DEV) / dim ** 0.5
    dz = torch.randn(M, r, device=DEV)
    dres = torch.randn(M, dim, device=DEV)
    # form 0: dense bf16 dy + additive rank-r term + residual, bf16 copy
    dy = torch.randn(M, dim, device=DEV).bfloat16()
    xr = x.double().requires_grad_(True)
    F.layer_norm(xr, (dim,), g.double(), be.double(), 1e-5).backward(dy.double())
    want = xr.grad + dres.double() + dz.double() @ w.double()
    dx_lp = torch.empty(M, dim, device=DEV, dtype=torch.bfloat16)
    n0 = ops.L.launch_count()
    got = ops.layernorm_bwd(x, g, mean, rstd, dy=dy, dres=dres, dx_lp=dx_lp, az=dz, aw=w, prec=ops.PREC_TF32)
    assert ops.L.launch_count() == n0 + 1
    close(got, want, 3e-3)
    close(dx_lp, want, 1e-2)
    exact = ops.layernorm_bwd(x, g, mean, rstd, dy=dy, dres=dres, az=dz, aw=w)
    close(exact, want, 1e-4)
    assert not torch.equal(exact, got)                      # the tf32 form really ran
    got2 = ops.layernorm_bwd(x, g, mean, rstd, dy=dy, az=dz, aw=w, prec=ops.PREC_TF32)      # no residual
    close(got2, want - dres.double(), 3e-3)
    # form 1: rank-r dy, residual in place, parameter gradients
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), be.double().requires_grad_(True)
    (F.layer_norm(xr, (dim,), gr, br, 1e-5) @ w.double().t()).backward(dz.double())
    dgam, dbet = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    buf = dres.clone()
    out = ops.layernorm_bwd(x, g, mean, rstd, dz=dz, w=w, dres=buf, dx=buf, dgamma=dgam, dbeta=dbet, prec=ops.PREC_TF32)
    assert out.data_ptr() == buf.data_ptr()
    close(out, xr.grad + dres.double(), 3e-3)
    close(dgam, gr.grad, 3e-3)
    close(dbet, br.grad, 3e-3)
    dgam2, dbet2 = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    exact = ops.layernorm_bwd(x, g, mean, rstd, dz=dz, w=w, dres=dres, dgamma=dgam2, dbeta=dbet2)
    close(exact, xr.grad + dres.double(), 1e-4)
    close(dgam2, gr.grad, 1e-4)


@pytest.mark.parametrize('M,dim,r', [(1033, 768, 20), (16 * 148 * 3 + 5, 768, 20), (517, 384, 20), (9, 768, 8), (16 * 148, 768, 24)])
def test_layernorm_bwd_with_output_projection(M, dim, r):
    """gvk_layernorm_bwd ow / oz: d(comb) = dG Wu computed from the output rows of the LayerNorm-backward pass that produces dG (in place over
    the residual gradient, bf16 copy written)."""
    torch.manual_seed(M + dim + r)
    x = torch.randn(M, dim, device=DEV) * 2 + 0.3
    g = torch.randn(dim, device=DEV) * 0.1 + 1
    be = torch.randn(dim, device=DEV) * 0.1
    _, mean, rstd = ops.layernorm_fwd(x, g, be)
    wu = torch.randn(dim, r, device=DEV) / dim ** 0.5          # nn.Linear(r, dim).weight
    dres = torch.randn(M, dim, device=DEV)
    dy = torch.randn(M, dim, device=DEV).bfloat16()
    xr = x.double().requires_grad_(True)
    F.layer_norm(xr, (dim,), g.double(), be.double(), 1e-5).backward(dy.double())
    want = xr.grad + dres.double()
    buf = dres.clone()
    dx_lp = torch.empty(M, dim, device=DEV, dtype=torch.bfloat16)
    n0 = ops.L.launch_count()
    dx, oz = ops.layernorm_bwd(x, g, mean, rstd, dy=dy, dres=buf, dx=buf, dx_lp=dx_lp, prec=ops.PREC_TF32, ow=wu, ow_transposed=True)
    assert ops.L.launch_count() == n0 + 1 and dx.data_ptr() == buf.data_ptr()
    close(dx, want, 1e-5)                                       # no rank term inside: the LayerNorm arithmetic is fp32
    close(dx_lp, want, 1e-2)
    close(oz, want @ wu.double(), 3e-3)
    close(oz, ops.rowproj_down(dx, wu, transposed=True, prec=ops.PREC_TF32)['z'], 3e-3)
    w2 = torch.randn(r, dim, device=DEV) / dim ** 0.5
    _, oz2 = ops.layernorm_bwd(x, g, mean, rstd, dy=dy, dres=dres, prec=ops.PREC_TF32, ow=w2)
    close(oz2, want @ w2.double().t(), 3e-3)
    with pytest.raises(Exception):                              # no exact-fp32 form of the fused projection
        ops.layernorm_bwd(x, g, mean, rstd, dy=dy, dres=dres, ow=w2)


@pytest.mark.parametrize('save', [True, False])
@pytest.mark.parametrize('M,dim,r', [(1033, 768, 20), (16 * 148 * 2 + 7, 768, 20), (517, 384, 20), (2066, 768, 32), (9, 768, 8)])
def test_layernorm_fwd_down_one_pass(M, dim, r, save):
    """gvk_layernorm_fwd_down: the LayerNorm output (bf16, + statistics) and the QuickGELU down-projection of the raw rows from one read of x."""
    torch.manual_seed(M + dim + r)
    x = torch.randn(M, dim, device=DEV) * 2 + 0.3
    g = torch.randn(dim, device=DEV) * 0.1 + 1
    be = torch.randn(dim, device=DEV) * 0.1
    w = torch.randn(r, dim, device=DEV) / dim ** 0.5
    b = torch.randn(r, device=DEV) * 0.1
    n0 = ops.L.launch_count()
    y, mean, rstd, d = ops.layernorm_fwd_down(x, g, be, w, b, act=ops.ROWACT_QUICKGELU, save_pre=save, save_stats=save)
    assert ops.L.launch_count() == n0 + 1
    yr = F.layer_norm(x.double(), (dim,), g.double(), be.double(), 1e-5)
    assert y.dtype == torch.bfloat16
    close(y, yr, 1e-2)
    y2, mean2, rstd2 = ops.layernorm_fwd(x, g, be, out_dtype=torch.bfloat16)
    assert (y.float() - y2.float()).abs().max().item() <= 2 ** -6 * yr.abs().max().item()      # at most a bf16 rounding step apart from the two-kernel path
    pre = x.double() @ w.double().t() + b.double()
    close(d['z'], pre * torch.sigmoid(1.702 * pre), 3e-3)
    if save:
        close(d['pre'], pre, 3e-3)
        close(mean, x.double().mean(1), 1e-5)
        close(rstd, 1.0 / (x.double().var(1, unbiased=False) + 1e-5).sqrt(), 1e-5)
    else:
        assert d['pre'] is None and mean is None and rstd is None
    # no bias, no activation
    _, _, _, d2 = ops.layernorm_fwd_down(x, g, be, w)
    close(d2['z'], x.double() @ w.double().t(), 3e-3)


@pytest.mark.parametrize('M,dim', [(1000, 768), (16 * 148 * 2 + 3, 768), (517, 384), (9, 768)])
def test_rowproj_up_down_one_pass(M, dim):
    """gvk_rowproj_up_down against the two kernels it replaces (same tf32 arithmetic, same dropout masks): forward form (bias, proj_drop,
    residual, QuickGELU down-projection) and backward form (in place over the residual, replayed mask in front of the down-projection)."""
    torch.manual_seed(M + dim)
    r, p = 20, 0.2
    c = torch.randn(M, r, device=DEV)
    wu = torch.randn(dim, r, device=DEV) / r ** 0.5          # nn.Linear(r, dim).weight
    bu = torch.randn(dim, device=DEV) * 0.1
    wd = torch.randn(r, dim, device=DEV) / dim ** 0.5        # nn.Linear(dim, r).weight
    bd = torch.randn(r, device=DEV) * 0.1
    res = torch.randn(M, dim, device=DEV)
    for drop in (0.0, p):
        n0 = ops.L.launch_count()
        out, d = ops.rowproj_up_down(c, wu, bu, res=res, up_drop_p=drop, up_seed=11, w2=wd, bias2=bd, act=ops.ROWACT_QUICKGELU, save_pre=True)
        assert ops.L.launch_count() == n0 + 1
        ref = ops.rowproj_up(c, wu, bu, res=res, drop_p=drop, seed=11, prec=ops.PREC_TF32)
        close(out, ref, 1e-6)                                  # same MMAs, same mask (an FMA contraction apart at most)
        assert torch.equal(out == res, ref == res)             # the same elements dropped
        rd = ops.rowproj_down(ref, wd, bd, act=ops.ROWACT_QUICKGELU, save_pre=True, prec=ops.PREC_TF32)
        close(d['pre'], rd['pre'], 1e-4)                       # same tf32 products, different summation order over the columns
        close(d['z'], rd['z'], 1e-4)
        if drop == 0.0:
            want = res.double() + c.double() @ wu.double().t() + bu.double()
            close(out, want, 3e-3)
            pre = want @ wd.double().t() + bd.double()
            close(d['pre'], pre, 3e-3)
            close(d['z'], pre * torch.sigmoid(1.702 * pre), 3e-3)
    # backward form: d(loc) += dul Wd in place, then dctx = mask(d(loc)) Wu with the forward's mask (seed 11)
    dul = torch.randn(M, r, device=DEV)
    dloc = torch.randn(M, dim, device=DEV)
    ref = ops.rowproj_up(dul, wd, transposed=True, res=dloc, prec=ops.PREC_TF32)
    rd = ops.rowproj_down(ref, wu, transposed=True, drop_p=p, seed=11, prec=ops.PREC_TF32)
    buf = dloc.clone()
    out, d = ops.rowproj_up_down(dul, wd, transposed=True, res=buf, out=buf, w2=wu, transposed2=True, dn_drop_p=p, dn_seed=11)
    assert out.data_ptr() == buf.data_ptr() and d['pre'] is None
    close(out, ref, 1e-6)
    close(d['z'], rd['z'], 1e-4)
    # no residual (first iteration of the backward loop)
    out0, d0 = ops.rowproj_up_down(dul, wd, transposed=True, w2=wu, transposed2=True)
    ref0 = ops.rowproj_up(dul, wd, transposed=True, prec=ops.PREC_TF32)
    close(out0, ref0, 1e-6)
    close(d0['z'], ops.rowproj_down(ref0, wu, transposed=True, prec=ops.PREC_TF32)['z'], 1e-4)


@pytest.mark.parametrize('prec', [ops.PREC_FP32, ops.PREC_TF32])
def test_rowproj_dropout_replay(prec):
    """The forward mask of rowproj_up is replayed by rowproj_down / skinny_wgrad in backward (same mask in both precisions)."""
    tol = 3e-3 if prec == ops.PREC_TF32 else 1e-4
    torch.manual_seed(3)
    M, dim, r, p = 640, 768, 20, 0.2
    c = torch.randn(M, r, device=DEV)
    wu = torch.randn(dim, r, device=DEV)
    ones = torch.ones(M, dim, device=DEV)
    # recover the mask: out = drop(c @ wu^T) with c @ wu^T replaced by a known tensor through r=... use bias-only path
    zero_c = torch.zeros(M, r, device=DEV)
    mask = ops.rowproj_up(zero_c, wu, torch.ones(dim, device=DEV), drop_p=p, seed=1234, prec=prec)        # = mask / (1-p)
    assert torch.equal(mask, ops.rowproj_up(zero_c, wu, torch.ones(dim, device=DEV), drop_p=p, seed=1234))
    keep = (mask > 0).float()
    assert abs(keep.mean().item() - (1 - p)) < 0.01
    assert torch.allclose(mask, keep / (1 - p))
    mask2 = ops.rowproj_up(zero_c, wu, torch.ones(dim, device=DEV), drop_p=p, seed=1235, prec=prec)
    assert (mask2 != mask).float().mean().item() > 0.2
    x = torch.randn(M, dim, device=DEV)
    close(ops.rowproj_down(x, wu, transposed=True, drop_p=p, seed=1234, prec=prec)['z'], (x * mask).double() @ wu.double(), tol)
    dw = torch.zeros(dim, r, device=DEV)
    dxc = torch.zeros(dim, device=DEV)
    ops.skinny_wgrad(c, x, dw=dw, dw_layout='dr', dx_colsum=dxc, drop_p=p, seed=1234, prec=prec)
    close(dw, (x * mask).double().t() @ c.double(), tol)
    close(dxc, (x * mask).double().sum(0), tol)
    del ones


def test_small_ops():
    torch.manual_seed(4)
    M = 3000
    a = torch.randn(M, 60, device=DEV)
    b = torch.randn(M, 20, device=DEV)
    dw = torch.zeros(60, 20, device=DEV)
    ops.small_wgrad(a, b, dw)
    close(dw, a.double().t() @ b.double(), 1e-4)
    w = torch.randn(60, 20, device=DEV)
    close(ops.small_matmul(a, w), a.double() @ w.double(), 1e-4)                     # row-per-thread form (60 -> 20)
    a96, w96 = torch.randn(M + 1, 96, device=DEV), torch.randn(96, 32, device=DEV)    # row-per-thread form (96 -> 32)
    close(ops.small_matmul(a96, w96), a96.double() @ w96.double(), 1e-4)
    a2, w2 = torch.randn(M + 3, 44, device=DEV), torch.randn(44, 12, device=DEV)      # generic form, odd sizes
    close(ops.small_matmul(a2, w2), a2.double() @ w2.double(), 1e-4)
    dw2 = torch.zeros(44, 12, device=DEV)
    ops.small_wgrad(a2, torch.randn(M + 3, 12, device=DEV) * 0 + 1, dw2)
    close(dw2, a2.double().sum(0)[:, None].expand(44, 12), 1e-4)
    a3, b3 = torch.randn(M + 3, 7, device=DEV), torch.randn(M + 3, 5, device=DEV)     # scalar staging path (sizes not multiples of 4)
    dw3 = torch.zeros(7, 5, device=DEV)
    ops.small_wgrad(a3, b3, dw3)
    close(dw3, a3.double().t() @ b3.double(), 1e-4)
    out = torch.zeros(60, device=DEV)
    ops.colsum(a, out)
    close(out, a.double().sum(0), 1e-4)
    # the hot-path shape: 64 000 rows, accumulation on top of an existing gradient (hi / lo split tf32 MMAs stay inside the fp32 tolerance)
    M4 = 64000
    a4, b4 = torch.randn(M4, 60, device=DEV), torch.randn(M4, 20, device=DEV)
    dw4 = torch.ones(60, 20, device=DEV)
    ops.small_wgrad(a4, b4, dw4)
    want4 = a4.double().t() @ b4.double() + 1.0
    assert (dw4.double() - want4).abs().max().item() <= 1e-5 * want4.abs().max().item()
    views = torch.randn(M, 100, device=DEV)                                           # strided operands
    dw5 = torch.zeros(60, 20, device=DEV)
    ops.small_wgrad(views[:, :60], views[:, 80:], dw5)
    close(dw5, views[:, :60].double().t() @ views[:, 80:].double(), 1e-4)
    x = torch.randn(77, 192, device=DEV)
    assert torch.equal(ops.cast_bf16(x), x.bfloat16())
    dy, pre = torch.randn(M, 20, device=DEV), torch.randn(M, 20, device=DEV)
    pr = pre.double().requires_grad_(True)
    (pr * torch.sigmoid(1.702 * pr)).backward(dy.double())
    close(ops.quickgelu_bwd(dy, pre), pr.grad)


@pytest.mark.parametrize('width', [64, 60, 72])
def test_split_pack_bf16_rule(width):
    """hi / lo bf16 slots of the K-extension operands: the 16-byte-store form (width % 8 == 0) and the element form write the same thing."""
    torch.manual_seed(width)
    rows, r = 1037, 20
    src = torch.randn(rows, r, device=DEV)
    big = torch.full((rows, 128 + width), 7.0, device=DEV, dtype=torch.bfloat16)
    for pattern in (0b010, 0b100):
        ops.split_pack_bf16(src, big[:, 128:], pattern)
        hi = src.bfloat16()
        lo = (src - hi.float()).bfloat16()
        want = torch.zeros(rows, width, device=DEV, dtype=torch.bfloat16)
        for slot in range(3):
            want[:, slot * r:(slot + 1) * r] = lo if (pattern >> slot) & 1 else hi
        assert torch.equal(big[:, 128:], want)
        assert (big[:, :128] == 7.0).all()


def _dense_attn_ref(qkv, B, T, H, D, scale, allow=None):
    dim = H * D
    q, k, v = qkv.double().view(B, T, 3, H, D).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) * scale
    if allow is not None:
        s = s.masked_fill(~allow.to(s.device), float('-inf'))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, dim)


@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
def test_attn_simt_dense(dt):
    torch.manual_seed(5)
    B, T, H, D = 2, 333, 3, 64
    dim = H * D
    qkv = torch.randn(B * T, 3 * dim, device=DEV).to(dt)
    o, lse = ops.attn_simt_fwd(qkv, B, T, H, D, q_off=0, k_off=dim, v_off=2 * dim, scale=D ** -0.5)
    qr = qkv.double().requires_grad_(True)
    ref = _dense_attn_ref(qr, B, T, H, D, D ** -0.5)
    tol = 3e-2 if dt == torch.bfloat16 else 1e-5
    close(o, ref.detach(), tol)
    do = torch.randn(B * T, dim, device=DEV).to(dt)
    ref.backward(do.double())
    dqkv = ops.attn_simt_bwd(qkv, o, lse, do, B, T, H, D, q_off=0, k_off=dim, v_off=2 * dim, scale=D ** -0.5)
    close(dqkv, qr.grad, 5e-2 if dt == torch.bfloat16 else 2e-5)


@pytest.mark.parametrize('prec', [ops.PREC_FP32, ops.PREC_TF32])
@pytest.mark.parametrize('dhw,k,r', [((10, 10, 10), (6, 6, 6), 20), ((4, 4, 4), (3, 2, 2), 20), ((5, 4, 3), (5, 4, 2), 20), ((3, 5, 7), (3, 6, 6), 20),
                                     ((10, 10, 10), (3, 6, 6), 32), ((2, 13, 13), (1, 4, 5), 20)])
def test_attn_simt_window(dhw, k, r, prec):
    """exact SIMT form and the tf32 tensor-core form (odd planes, chunks of several planes, planes larger than a chunk, r = 32)"""
    torch.manual_seed(6)
    B = 2
    N = dhw[0] * dhw[1] * dhw[2]
    qkv = torch.randn(B * N, 3 * r, device=DEV)
    allow = O.window_allow(dhw, k)
    tol = 1e-5 if prec == ops.PREC_FP32 else 2e-3
    kw = dict(q_off=0, k_off=r, v_off=2 * r, scale=0.3, window=k, grid=dhw, prec=prec)
    o, lse = ops.attn_simt_fwd(qkv, B, N, 1, r, **kw)
    qr = qkv.double().requires_grad_(True)
    ref = _dense_attn_ref(qr, B, N, 1, r, 0.3, allow)
    close(o, ref.detach(), tol)
    s = torch.einsum('bid,bjd->bij', qr.detach().view(B, N, 3 * r)[..., :r], qr.detach().view(B, N, 3 * r)[..., r:2 * r]) * 0.3
    close(lse, torch.logsumexp(s.masked_fill(~allow.to(DEV), float('-inf')), -1).reshape(-1), tol)
    do = torch.randn(B * N, r, device=DEV)
    ref.backward(do.double())
    dqkv = ops.attn_simt_bwd(qkv, o, lse, do, B, N, 1, r, **kw)
    close(dqkv, qr.grad, 2 * tol)
    for blk in range(3):
        close(dqkv[:, blk * r:(blk + 1) * r], qr.grad[:, blk * r:(blk + 1) * r], 2 * tol)


@pytest.mark.parametrize('prec', [ops.PREC_FP32, ops.PREC_TF32])
def test_attn_window_two_heads(prec):
    """head h of q / k / v at columns q_off / k_off / v_off + h * D (the API allows H > 1 although LocalSelfAttention uses one head)"""
    torch.manual_seed(9)
    B, H, r, dhw, k = 2, 2, 20, (4, 5, 6), (3, 4, 4)
    N = dhw[0] * dhw[1] * dhw[2]
    qkv = torch.randn(B * N, 3 * H * r, device=DEV)
    allow = O.window_allow(dhw, k)
    kw = dict(q_off=0, k_off=H * r, v_off=2 * H * r, scale=0.3, window=k, grid=dhw, prec=prec)
    tol = 1e-5 if prec == ops.PREC_FP32 else 2e-3
    o, lse = ops.attn_simt_fwd(qkv, B, N, H, r, **kw)
    qr = qkv.double().requires_grad_(True)
    ref = _dense_attn_ref(qr, B, N, H, r, 0.3, allow)
    close(o, ref.detach(), tol)
    do = torch.randn(B * N, H * r, device=DEV)
    ref.backward(do.double())
    close(ops.attn_simt_bwd(qkv, o, lse, do, B, N, H, r, **kw), qr.grad, 2 * tol)


@pytest.mark.parametrize('prec', [ops.PREC_FP32, ops.PREC_TF32])
def test_attn_window_dropout_statistics_and_replay(prec):
    torch.manual_seed(7)
    B, r, dhw, k, p = 2, 20, (10, 10, 10), (6, 6, 6), 0.2
    N = 1000
    qkv = torch.randn(B * N, 3 * r, device=DEV)
    qkv[:, 2 * r:] = 1.0                      # v == 1  =>  out_i = sum_j p_ij m_ij / (1-p), expectation 1
    o, lse = ops.attn_simt_fwd(qkv, B, N, 1, r, q_off=0, k_off=r, v_off=2 * r, scale=0.05, window=k, grid=dhw, drop_p=p, seed=99, prec=prec)
    assert abs(o.mean().item() - 1.0) < 0.01
    assert o.std().item() > 0.01              # masks are actually applied
    o2, _ = ops.attn_simt_fwd(qkv, B, N, 1, r, q_off=0, k_off=r, v_off=2 * r, scale=0.05, window=k, grid=dhw, drop_p=p, seed=99, prec=prec)
    assert torch.equal(o, o2)                 # replayable


@pytest.mark.parametrize('prec', [ops.PREC_FP32, ops.PREC_TF32])
def test_attn_window_dropout_mask_is_the_same_in_forward_and_backward(prec):
    """Recover the dropout mask from the forward (q = k = 0 gives uniform probabilities, V = indicator columns), then check the forward and
    every gradient block against autograd through a dense reference that applies that mask."""
    torch.manual_seed(8)
    B, r, dhw, k, p = 2, 20, (4, 5, 6), (3, 4, 4), 0.25
    N = dhw[0] * dhw[1] * dhw[2]
    allow = O.window_allow(dhw, k).to(DEV)
    kw = dict(q_off=0, k_off=r, v_off=2 * r, scale=0.3, window=k, grid=dhw, drop_p=p, seed=11, offset=77, prec=prec)
    mask = torch.zeros(B, N, N, device=DEV, dtype=torch.float64)
    for c0 in range(0, N, r):
        probe = torch.zeros(B, N, 3 * r, device=DEV)
        probe[:, c0:c0 + r, 2 * r:] = torch.eye(r, device=DEV)
        o, _ = ops.attn_simt_fwd(probe.view(B * N, 3 * r), B, N, 1, r, **kw)
        mask[:, :, c0:c0 + r] = o.view(B, N, r).double() * allow.sum(1).double()[None, :, None] * (1 - p)
    assert ((mask - mask.round()).abs() < 5e-3).all()
    mask = mask.round()
    assert set(mask.unique().tolist()) == {0.0, 1.0} and (mask[~allow.expand(B, N, N)] == 0).all()
    keep = mask[allow.expand(B, N, N)].mean().item()
    assert abs(keep - (1 - p)) < 0.02 and not torch.equal(mask[0], mask[1])
    qkv = torch.randn(B * N, 3 * r, device=DEV)
    do = torch.randn(B * N, r, device=DEV)
    qr = qkv.double().requires_grad_(True)
    q, kk, v = qr.view(B, N, 3, r).unbind(2)
    s = (q @ kk.transpose(1, 2) * 0.3).masked_fill(~allow, float('-inf'))
    ref = ((torch.softmax(s, -1) * mask / (1 - p)) @ v).reshape(B * N, r)
    ref.backward(do.double())
    tol = 1e-5 if prec == ops.PREC_FP32 else 2e-3
    o, lse = ops.attn_simt_fwd(qkv, B, N, 1, r, **kw)
    close(o, ref.detach(), tol)
    dqkv = ops.attn_simt_bwd(qkv, o, lse, do, B, N, 1, r, **kw)
    for blk in range(3):
        close(dqkv[:, blk * r:(blk + 1) * r], qr.grad[:, blk * r:(blk + 1) * r], 2 * tol)


def _fusion_sd(P, r, dim, dev):
    g = torch.Generator().manual_seed(11)
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    sd = {'proj_down.0.weight': rn(r, dim) / dim ** 0.5, 'proj_down.0.bias': rn(r) * 0.1, 'proj_up.weight': rn(dim, r) / r ** 0.5, 'proj_up.bias': rn(dim) * 0.1,
          'cls_analyzer.cls_analyzer_.0.weight': 1 + 0.1 * rn(r), 'cls_analyzer.cls_analyzer_.0.bias': 0.1 * rn(r),
          'cls_analyzer.cls_analyzer_.1.weight': rn(64, r) / r ** 0.5, 'cls_analyzer.cls_analyzer_.1.bias': 0.1 * rn(64),
          'cls_analyzer.cls_analyzer_.3.weight': rn(P, 64) / 8, 'cls_analyzer.cls_analyzer_.3.bias': 0.1 * rn(P),
          'gl_balancer.gl_balancer_.0.weight': 1 + 0.1 * rn(r), 'gl_balancer.gl_balancer_.0.bias': 0.1 * rn(r),
          'gl_balancer.gl_balancer_.1.weight': rn(1, r) / r ** 0.5, 'gl_balancer.gl_balancer_.1.bias': 0.1 * rn(1),
          'global_attention.query_proj.weight': rn(r, r) / r ** 0.5, 'global_attention.query_proj.bias': 0.1 * rn(r),
          'local_attention.query_proj.weight': rn(r, r) / r ** 0.5, 'local_attention.query_proj.bias': 0.1 * rn(r)}
    return {k: v.to(dev) for k, v in sd.items()}


FUSION_KEYS = {'wq_g': 'global_attention.query_proj.weight', 'bq_g': 'global_attention.query_proj.bias', 'wq_l': 'local_attention.query_proj.weight',
               'bq_l': 'local_attention.query_proj.bias', 'a_ln_w': 'cls_analyzer.cls_analyzer_.0.weight', 'a_ln_b': 'cls_analyzer.cls_analyzer_.0.bias',
               'a_w1': 'cls_analyzer.cls_analyzer_.1.weight', 'a_b1': 'cls_analyzer.cls_analyzer_.1.bias', 'a_w3': 'cls_analyzer.cls_analyzer_.3.weight',
               'a_b3': 'cls_analyzer.cls_analyzer_.3.bias', 'g_ln_w': 'gl_balancer.gl_balancer_.0.weight', 'g_ln_b': 'gl_balancer.gl_balancer_.0.bias',
               'g_w': 'gl_balancer.gl_balancer_.1.weight', 'g_b': 'gl_balancer.gl_balancer_.1.bias'}


@pytest.mark.parametrize('P,N,r', [(32, 1000, 20), (8, 64, 20), (40, 333, 20), (5, 97, 32), (33, 50, 16)])
def test_prompt_fusion_fwd_bwd(P, N, r):
    """Whole Awakening_Prompt (down-proj, fusion core, up-proj) against the oracle restatement in fp64 (more than 32 prompts = two lane
    groups per CTA, latent widths 16 / 20 / 32, key sets that do / do not fit the shared-memory staging)."""
    torch.manual_seed(12)
    B, dim = 2, 192
    T = P + 1 + N
    sd = _fusion_sd(P, r, dim, DEV)
    g = torch.randn(B, T, dim, device=DEV)
    loc = torch.randn(B, N, dim, device=DEV)
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    g64, loc64 = g.double().requires_grad_(True), loc.double().requires_grad_(True)
    ref = O.awakening_prompt(g64, loc64, sd64, '', P)
    dprompt = torch.randn(B, T, dim, device=DEV)
    ref.backward(dprompt.double())
    # ours
    k = {kk: sd[v].contiguous() for kk, v in FUSION_KEYS.items()}
    dg = ops.rowproj_down(g.view(B * T, dim), sd['proj_down.0.weight'], sd['proj_down.0.bias'], act=ops.ROWACT_QUICKGELU, save_pre=True)
    dl = ops.rowproj_down(loc.view(B * N, dim), sd['proj_down.0.weight'], sd['proj_down.0.bias'], act=ops.ROWACT_QUICKGELU, save_pre=True)
    comb, ll = dg['z'], dl['z']
    saved = ops.prompt_fusion_fwd(comb, ll, k, B, T, N, P)
    out = ops.rowproj_up(comb, sd['proj_up.weight'], sd['proj_up.bias'])
    close(out, ref.detach().view(B * T, dim), 2e-5)
    dG = dprompt.view(B * T, dim).contiguous()
    dcomb = ops.rowproj_down(dG, sd['proj_up.weight'], transposed=True)['z']
    grads = {kk: torch.zeros_like(v) for kk, v in k.items()}
    dll = ops.prompt_fusion_bwd(comb, ll, dcomb, k, saved, grads, B, T, N, P)
    du = ops.quickgelu_bwd(dcomb, dg['pre'])
    dul = ops.quickgelu_bwd(dll, dl['pre'])
    dg_in = ops.rowproj_up(du, sd['proj_down.0.weight'], transposed=True)
    dloc_in = ops.rowproj_up(dul, sd['proj_down.0.weight'], transposed=True)
    close(dg_in, g64.grad.view(B * T, dim), 1e-4)
    close(dloc_in, loc64.grad.view(B * N, dim), 1e-4)
    for kk, name in FUSION_KEYS.items():
        close(grads[kk], sd64[name].grad, 2e-4)
    dwd = torch.zeros(r, dim, device=DEV)
    dbd = torch.zeros(r, device=DEV)
    ops.skinny_wgrad(du, g.view(B * T, dim), dw=dwd, da_colsum=dbd)
    ops.skinny_wgrad(dul, loc.view(B * N, dim), dw=dwd, da_colsum=dbd)
    close(dwd, sd64['proj_down.0.weight'].grad, 2e-4)
    close(dbd, sd64['proj_down.0.bias'].grad, 2e-4)
    dwu = torch.zeros(dim, r, device=DEV)
    dbu = torch.zeros(dim, device=DEV)
    ops.skinny_wgrad(comb, dG, dw=dwu, dw_layout='dr', dx_colsum=dbu)
    close(dwu, sd64['proj_up.weight'].grad, 2e-4)
    close(dbu, sd64['proj_up.bias'].grad, 2e-4)


@pytest.mark.parametrize('pool', [(0, 33), (0, 1), (0, 1033)])
def test_head_fwd_bwd(pool):
    torch.manual_seed(13)
    B, T, dim, C = 3, 1033, 192, 5
    x = torch.randn(B * T, dim, device=DEV)
    g, b = 1 + 0.1 * torch.randn(dim, device=DEV), 0.1 * torch.randn(dim, device=DEV)
    wh, bh = torch.randn(C, dim, device=DEV) / dim ** 0.5, 0.1 * torch.randn(C, device=DEV)
    logits, pooled = ops.head_fwd(x, B, T, pool[0], pool[1], g, b, wh, bh)
    xr = x.double().requires_grad_(True)
    whr, bhr = wh.double().requires_grad_(True), bh.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    y = F.layer_norm(xr.view(B, T, dim), (dim,), gr, br, 1e-5)[:, pool[0]:pool[0] + pool[1]].mean(1)
    ref = y @ whr.t() + bhr
    close(logits, ref.detach())
    dl = torch.randn(B, C, device=DEV)
    ref.backward(dl.double())
    dgam, dbet = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    dx, dwh, dbh = ops.head_bwd(x, B, T, pool[0], pool[1], g, b, wh, bh, pooled, dl, dgamma=dgam, dbeta=dbet)
    close(dx, xr.grad, 2e-5)
    close(dwh, whr.grad)
    close(dbh, bhr.grad)
    close(dgam, gr.grad, 1e-4)
    close(dbet, br.grad, 1e-4)


def test_losses_against_oracle_and_known_answers():
    from helpers import load_golden
    g = load_golden('focal_known_answers')
    for i in range(4):
        z = torch.tensor(g[f'z{i}'], device=DEV)
        y = torch.tensor(g[f'y{i}'], device=DEV)
        loss, dz = ops.loss_fwd_bwd(z, y, ops.LOSS_FOCAL)
        assert loss.item() == pytest.approx(float(g[f'loss{i}']), rel=2e-6)
        assert torch.allclose(dz.cpu(), torch.tensor(g[f'dz{i}']), rtol=2e-5, atol=1e-8)
    torch.manual_seed(14)
    z = torch.randn(37, 5, device=DEV)
    y = torch.randint(0, 5, (37,), device=DEV)
    loss, dz = ops.loss_fwd_bwd(z, y, ops.LOSS_CE)
    zr = z.double().requires_grad_(True)
    ref = F.cross_entropy(zr, y)
    ref.backward()
    assert loss.item() == pytest.approx(ref.item(), rel=1e-6)
    close(dz, zr.grad)


def test_patch_gather_and_token_rows():
    torch.manual_seed(15)
    B, D, H, W, fp, ps = 2, 24, 32, 48, 12, 16
    img = torch.rand(B, 1, D, H, W, device=DEV)
    ref = O.patchify(img, fp, ps)
    got = ops.patch_gather(img, fp, ps, torch.float32)
    assert torch.equal(got.view_as(ref), ref)
    assert torch.equal(ops.patch_gather(img, fp, ps, torch.bfloat16).view_as(ref), ref.bfloat16())
    a, b = torch.randn(5, 192, device=DEV), torch.randn(5, 192, device=DEV)
    out = torch.zeros(3 * 20, 192, device=DEV)
    ops.fill_rows(a, b, out, 20, 2, 3)
    assert torch.equal(out.view(3, 20, 192)[:, 2:7], (a + b).expand(3, -1, -1))
    assert out.view(3, 20, 192)[:, 7:].abs().max().item() == 0
    x = torch.randn(3 * 20, 192, device=DEV)
    close(ops.batch_rowsum(x, 20, 2, 5, 3), x.view(3, 20, 192)[:, 2:7].double().sum(0))


# ---------------------------------------------------------------------------------------------- DVPT side path (csrc/gvk_dvpt.cu)
@pytest.mark.parametrize('B,T,P', [(3, 69, 4), (2, 71, 6), (2, 1051, 50), (1, 1033, 32)])
def test_dvpt_latent_xattn_fwd_bwd(B, T, P):
    """Prompt -> token cross attention in the 20-wide latent (model/dvpt.py:38-45) against its torch restatement (tests/ops_double.py)."""
    import ops_double as D
    torch.manual_seed(T + P)
    r, scale = 20, 192 ** -0.5
    z = torch.randn(B * T, r, device=DEV)
    zc, zr = z.clone(), z.cpu().clone()
    pl, lse = ops.latent_xattn_fwd(zc, B, T, P, scale)
    plr, lser = D.latent_xattn_fwd(zr, B, T, P, scale)
    close(zc.cpu(), zr)
    close(pl.cpu(), plr)
    close(lse.cpu(), lser)
    dz = torch.randn(B * T, r, device=DEV)
    dzr = dz.cpu().clone()
    ops.latent_xattn_bwd(zc, pl, lse, dz, B, T, P, scale)
    D.latent_xattn_bwd(zr, plr, lser, dzr, B, T, P, scale)
    close(dz.cpu(), dzr, 1e-4)


def test_dvpt_quickgelu_and_gate_kernels():
    import ops_double as D
    torch.manual_seed(3)
    x = torch.randn(517, 768, device=DEV) * 2
    close(ops.quickgelu_fwd(x).cpu(), D.quickgelu_fwd(x.cpu()))
    dy, res = torch.randn_like(x), torch.randn_like(x)
    close(ops.quickgelu_bwd_add(dy, x, res=res).cpu(), D.quickgelu_bwd_add(dy.cpu(), x.cpu(), res=res.cpu()))
    close(ops.quickgelu_bwd_add(dy, x).cpu(), D.quickgelu_bwd_add(dy.cpu(), x.cpu()))
    out = res.clone()
    ops.quickgelu_bwd_add(dy, x, res=out, out=out)                 # in place over the residual
    close(out.cpu(), D.quickgelu_bwd_add(dy.cpu(), x.cpu(), res=res.cpu()))
    w, dw = torch.randn(768, 20, device=DEV), torch.randn(768, 20, device=DEV)
    gate = torch.tensor([0.37], device=DEV)
    close(ops.gate_scale(w, gate).cpu(), 0.37 * w.cpu())
    acc, dgate = torch.ones_like(w), torch.tensor([2.0], device=DEV)
    ops.gate_grads(w, dw, gate, acc, dgate)
    close(acc.cpu(), 1 + 0.37 * dw.cpu())
    close(dgate.cpu(), 2.0 + (w.double() * dw.double()).sum().cpu().reshape(1), 1e-5)


# ---------------------------------------------------------------------------------------------- EVP primitives (csrc/gvk_evp.cu)
@pytest.mark.parametrize('M,na,nb', [(1000, 192, 64), (2050, 768, 192), (515, 64, 3072), (33, 4, 8), (4099, 132, 68)])
@pytest.mark.parametrize('dta,dtb', [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)])
def test_evp_wgrad_general(M, na, nb, dta, dtb):
    """dw += a^T b over the rows (autograd's Linear / Conv3d weight gradients of model/evp.py:42-52) against fp64: exact mode to fp32 rounding,
    tf32 mode to the tf32 operand rounding; dw accumulates; ragged edges on every side of the 128 x 64 x 32 tiles."""
    torch.manual_seed(M + na)
    a = torch.randn(M, na, device=DEV).to(dta)
    b = torch.randn(M, nb, device=DEV).to(dtb)
    ref = a.double().t() @ b.double()
    scale = ref.abs().max().item()
    base = torch.randn(na, nb, device=DEV)
    for prec, tol in ((ops.PREC_FP32, 2e-6), (ops.PREC_TF32, 2e-3)):
        dw = base.clone()
        ops.wgrad(a, b, dw, prec=prec)
        err = (dw.double() - base.double() - ref).abs().max().item()
        assert err <= tol * scale, (prec, err, scale)
    # accumulate into a column slice of a wider matrix (ld_dw > nb)
    wide = torch.zeros(na, nb + 8, device=DEV)
    ops.wgrad(a, b, wide[:, 4:4 + nb], prec=ops.PREC_FP32)
    assert wide[:, :4].abs().max().item() == 0 and wide[:, 4 + nb:].abs().max().item() == 0
    assert (wide[:, 4:4 + nb].double() - ref).abs().max().item() <= 2e-6 * scale


def test_evp_wgrad_row_maps():
    """Operands that skip the cls row of every volume (row map (rows_per_batch, batch_rows) on either side) without a copy."""
    torch.manual_seed(5)
    B, N, T, na, nb = 3, 64, 65, 64, 192
    a_full = torch.randn(B * T, na, device=DEV)
    b_c = torch.randn(B * N, nb, device=DEV).bfloat16()
    a_tok = a_full.view(B, T, na)[:, 1:].reshape(B * N, na)
    ref = a_tok.double().t() @ b_c.double()
    dw = torch.zeros(na, nb, device=DEV)
    ops.wgrad(a_full[1:], b_c, dw, M=B * N, a_rows=(N, T))
    assert (dw.double() - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    dw2 = torch.zeros(nb, na, device=DEV)
    ops.wgrad(b_c, a_full[1:], dw2, M=B * N, b_rows=(N, T))
    assert (dw2.double() - ref.t()).abs().max().item() <= 2e-6 * ref.abs().max().item()


@pytest.mark.parametrize('shape,rate', [((2, 1, 48, 64, 64), 0.25), ((1, 1, 24, 32, 48), 0.1), ((2, 1, 12, 32, 32), 0.9), ((1, 2, 20, 16, 24), 0.5),
                                        ((2, 1, 120, 160, 160), 0.25), ((1, 1, 6, 200, 330), 0.3), ((1100, 1, 60, 8, 12), 0.25)])
def test_evp_hfreq_filter_matches_reference_fft(shape, rate):
    """gvk_hfreq_filter with the engine's closed-form filter against PromptGenerator.fft restated with torch.fft (oracle.evp_highpass, pinned
    to the live reference by the evp goldens): fftshift over every axis, mask on the depth / height axes, real part, abs."""
    from gaviko_b200.model.evp import ExplicitVisualPrompting
    from gaviko_b200.vit_engine import VitEngine

    class _Stub:      # the filter builder only reads prompt_generator.freq_nums
        class prompt_generator:
            freq_nums = rate
    eng = VitEngine.__new__(VitEngine)
    eng.__dict__['_module_ref'] = _Stub
    img = torch.rand(shape, generator=torch.Generator().manual_seed(7)).to(DEV)
    filt, hit = eng._evp_filter(img)
    out = ops.hfreq_filter(img, filt, hit)
    ref = O.evp_highpass(img.cpu(), rate)
    assert (out.cpu() - ref).abs().max().item() < 5e-6
    assert ExplicitVisualPrompting is not None


@pytest.mark.parametrize('dt_in,dt_out,with_res', [(torch.float32, torch.float32, False), (torch.bfloat16, torch.bfloat16, False), (torch.float32, torch.float32, True),
                                                   (torch.bfloat16, torch.float32, True)])
def test_elementwise_dropout_mask_rule_and_views(dt_in, dt_out, with_res):
    """gvk_dropout (the nn.Dropout sites of model/vision_transformer.py:33,35,58,157 in the un-frozen train mode): out = res + x * mask with the mask
    rule of include/gvk.h restated on the host (helpers.elementwise_keep_mask): 16-element groups (one Philox call each), the 4-element path
    (N or offset not a multiple of 16: same rule, same mask), contiguous tensors (vector accesses) and odd-offset column views (scalar accesses)."""
    from helpers import elementwise_keep_mask
    torch.manual_seed(11)
    p, seed = 0.3, 0x1234567890ABCDEF & ((1 << 63) - 1)
    for M, N, off in ((203, 192, 16), (203, 192, 8), (57, 72, 4), (64, 3072, 0)):
        x = torch.randn(M, N, device=DEV).to(dt_in)
        res = torch.randn(M, N, device=DEV) if with_res else None
        keep, scale = elementwise_keep_mask(M, N, p, seed, off)
        keep = keep.to(DEV)
        tol = 1e-6 if dt_out == torch.float32 else 1e-2
        ref = x.float() * keep * scale + (res if with_res else 0)
        out = ops.dropout(x, p, seed, res=res, out_dtype=dt_out, offset=off)
        assert out.dtype == dt_out
        close(out, ref.to(dt_out), tol)
        assert abs(keep.float().mean().item() - (1 - p)) < 0.03
        # unaligned views: columns 1 .. of a wider input, columns 3 .. of a wider output (pointers off the 16-byte grid, odd leading dimensions)
        wide = torch.randn(M, N + 5, device=DEV).to(dt_in)
        xv = wide[:, 1:1 + N]
        out2 = torch.empty(M, N + 3, device=DEV, dtype=dt_out)[:, 3:3 + N]
        ops.dropout(xv, p, seed, res=res, out=out2, offset=off)
        close(out2, (xv.float() * keep * scale + (res if with_res else 0)).to(dt_out), tol)
    # replay: the same (seed, offset) gives the same mask, another seed another one
    a = ops.dropout(torch.ones(64, 768, device=DEV), p, seed)
    assert torch.equal(a, ops.dropout(torch.ones(64, 768, device=DEV), p, seed)) and not torch.equal(a, ops.dropout(torch.ones(64, 768, device=DEV), p, seed + 1))
