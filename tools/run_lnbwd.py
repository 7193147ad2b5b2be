"""Time the LayerNorm-backward forms of the GAViKO step that carry a rank-20 product (d(g_mid) pass: dense bf16 dy + az @ aw; local-branch
pass: dy = dz @ w with dgamma / dbeta) — exact-fp32 kernel vs the tensor-core kernels — against the HBM roofline.
python tools/run_lnbwd.py [--batch 64]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gaviko_b200 import ops  # noqa: E402
from run_rowops import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--flush', default='dirty', choices=['dirty', 'clean', 'none'])
    ap.add_argument('--only', default='', help="'updown': only the up + down pair")
    a = ap.parse_args()
    dim, r, peak = 768, 20, 6550.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:  # noqa: BLE001
        pass
    dev = 'cuda'
    torch.manual_seed(0)
    if a.only == 'updown':
        return updown_pair(a, peak)
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    gamma = torch.rand(dim, device=dev) + 0.5
    w = torch.randn(r, dim, device=dev) / dim ** 0.5
    rows = []
    for name, M in (('g_mid pass (bf16 dy + dres + az @ aw -> dx fp32 + bf16)', a.batch * 1033), ('local pass (dz @ w + dres in place + dgamma / dbeta)', a.batch * 1000)):
        x = torch.randn(M, dim, device=dev)
        mean, rstd = x.mean(1), 1.0 / x.var(1, unbiased=False).add(1e-5).sqrt()
        dz = torch.randn(M, r, device=dev)
        dres = torch.randn(M, dim, device=dev)
        if name.startswith('g_mid'):
            dy = torch.randn(M, dim, device=dev).bfloat16()
            dx = torch.empty(M, dim, device=dev)
            dx_lp = torch.empty(M, dim, device=dev, dtype=torch.bfloat16)
            nbytes = M * dim * (4 + 2 + 4 + 4 + 2)
            forms = {'exact fp32': ops.PREC_FP32, 'tensor-core (tf32 mma.sync, cp.async-staged streams)': ops.PREC_TF32}
            def run(prec):
                ops.layernorm_bwd(x, gamma, mean, rstd, dy=dy, dres=dres, dx=dx, dx_lp=dx_lp, az=dz, aw=w, prec=prec)
        else:
            dgam, dbet = torch.zeros(dim, device=dev), torch.zeros(dim, device=dev)
            nbytes = M * dim * (4 + 4 + 4)
            forms = {'exact fp32': ops.PREC_FP32, 'tensor-core (tf32 mma.sync, cp.async-staged streams)': ops.PREC_TF32}
            def run(prec):
                ops.layernorm_bwd(x, gamma, mean, rstd, dz=dz, w=w, dres=dres, dx=dres, dgamma=dgam, dbeta=dbet, prec=prec)
        for form, prec in forms.items():
            us = timeit(lambda: run(prec), a.iters, flush, a.flush)
            rows.append(dict(kernel=name, form=form, M=M, us=round(us, 1), gbps=round(nbytes / us / 1e3, 1), frac_of_copy_peak=round(nbytes / us / 1e3 / peak, 3),
                             roofline_us=round(nbytes / peak / 1e3, 1)))
            print(json.dumps(rows[-1]), flush=True)
    forward_pair(a, peak)
    backward_pair(a, peak)
    updown_pair(a, peak)


def backward_pair(a, peak):
    """LayerNorm1 backward producing dG (in place over d(g_mid), + bf16 copy) and d(comb) = dG Wu of the next iteration: two kernels vs one."""
    dev, dim, r = 'cuda', 768, 20
    M = a.batch * 1033
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    x = torch.randn(M, dim, device=dev)
    gamma = torch.rand(dim, device=dev) + 0.5
    mean, rstd = x.mean(1), 1.0 / x.var(1, unbiased=False).add(1e-5).sqrt()
    dy = torch.randn(M, dim, device=dev).bfloat16()
    dres = torch.randn(M, dim, device=dev)
    dx_lp = torch.empty(M, dim, device=dev, dtype=torch.bfloat16)
    wu = torch.randn(dim, r, device=dev) / dim ** 0.5
    nb = M * dim * (4 + 2 + 4 + 4 + 2)
    forms = {
        'layernorm_bwd exact (dy bf16 + dres in place + bf16 copy)': (lambda: ops.layernorm_bwd(x, gamma, mean, rstd, dy=dy, dres=dres, dx=dres, dx_lp=dx_lp), nb),
        'rowproj_down (tf32, transposed weight)': (lambda: ops.rowproj_down(dres, wu, transposed=True, prec=ops.PREC_TF32), M * dim * 4),
        'layernorm_bwd with the output projection (one pass)': (lambda: ops.layernorm_bwd(x, gamma, mean, rstd, dy=dy, dres=dres, dx=dres, dx_lp=dx_lp, prec=ops.PREC_TF32,
                                                                                       ow=wu, ow_transposed=True), nb),
    }
    for form, (fn, nbytes) in forms.items():
        us = timeit(fn, a.iters, flush, a.flush)
        print(json.dumps(dict(kernel='dG pass + next d(comb)', form=form, M=M, us=round(us, 1), gbps=round(nbytes / us / 1e3, 1),
                              frac_of_copy_peak=round(nbytes / us / 1e3 / peak, 3), roofline_us=round(nbytes / peak / 1e3, 1))), flush=True)


def updown_pair(a, peak):
    """Local branch: proj_up + proj_drop + residual then Awakening_Prompt.proj_down (forward), d(loc) += d(ul) Wd then the dgrad of proj_up with the
    replayed mask (backward): two kernels vs gvk_rowproj_up_down."""
    dev, dim, r, p = 'cuda', 768, 20, 0.2
    M = a.batch * 1000
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    c = torch.randn(M, r, device=dev)
    wu, bu = torch.randn(dim, r, device=dev) / r ** 0.5, torch.randn(dim, device=dev) * 0.1
    wd, bd = torch.randn(r, dim, device=dev) / dim ** 0.5, torch.randn(r, device=dev) * 0.1
    res, out = torch.randn(M, dim, device=dev), torch.empty(M, dim, device=dev)
    nb = M * dim * 8
    forms = {
        'fwd: rowproj_up (dropout 0.2)': (lambda: ops.rowproj_up(c, wu, bu, res=res, out=out, drop_p=p, seed=3, prec=ops.PREC_TF32), nb),
        'fwd: rowproj_down (QuickGELU, pre saved)': (lambda: ops.rowproj_down(out, wd, bd, act=ops.ROWACT_QUICKGELU, save_pre=True, prec=ops.PREC_TF32), nb // 2),
        'fwd: rowproj_up_down (dropout 0.2)': (lambda: ops.rowproj_up_down(c, wu, bu, res=res, out=out, up_drop_p=p, up_seed=3, w2=wd, bias2=bd, act=ops.ROWACT_QUICKGELU,
                                                                          save_pre=True), nb),
        'fwd: rowproj_up_down (no dropout: inference)': (lambda: ops.rowproj_up_down(c, wu, bu, res=res, out=out, w2=wd, bias2=bd, act=ops.ROWACT_QUICKGELU), nb),
        'bwd: rowproj_up (in place)': (lambda: ops.rowproj_up(c, wd, transposed=True, res=res, out=res, prec=ops.PREC_TF32), nb),
        'bwd: rowproj_down (replayed mask)': (lambda: ops.rowproj_down(res, wu, transposed=True, drop_p=p, seed=3, prec=ops.PREC_TF32), nb // 2),
        'bwd: rowproj_up_down (in place, replayed mask)': (lambda: ops.rowproj_up_down(c, wd, transposed=True, res=res, out=res, w2=wu, transposed2=True, dn_drop_p=p, dn_seed=3), nb),
    }
    for form, (fn, nbytes) in forms.items():
        us = timeit(fn, a.iters, flush, a.flush)
        print(json.dumps(dict(kernel='local stream up + down', form=form, M=M, us=round(us, 1), gbps=round(nbytes / us / 1e3, 1),
                              frac_of_copy_peak=round(nbytes / us / 1e3 / peak, 3), roofline_us=round(nbytes / peak / 1e3, 1))), flush=True)


def forward_pair(a, peak):
    """FeedForward's LayerNorm + Awakening_Prompt.proj_down on g_mid: two kernels vs gvk_layernorm_fwd_down."""
    dev, dim, r = 'cuda', 768, 20
    M = a.batch * 1033
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    x = torch.randn(M, dim, device=dev)
    gamma, beta = torch.rand(dim, device=dev) + 0.5, torch.randn(dim, device=dev) * 0.1
    w, b = torch.randn(r, dim, device=dev) / dim ** 0.5, torch.randn(r, device=dev) * 0.1
    y = torch.empty(M, dim, device=dev, dtype=torch.bfloat16)
    forms = {
        'layernorm_fwd (bf16 out)': (lambda: ops.layernorm_fwd(x, gamma, beta, out=y), M * dim * 6),
        'rowproj_down (tf32, QuickGELU, pre saved)': (lambda: ops.rowproj_down(x, w, b, act=ops.ROWACT_QUICKGELU, save_pre=True, prec=ops.PREC_TF32), M * dim * 4),
        'layernorm_fwd_down (one pass)': (lambda: ops.layernorm_fwd_down(x, gamma, beta, w, b, act=ops.ROWACT_QUICKGELU, save_pre=True), M * dim * 6),
    }
    for form, (fn, nbytes) in forms.items():
        us = timeit(fn, a.iters, flush, a.flush)
        print(json.dumps(dict(kernel='g_mid forward readers', form=form, M=M, us=round(us, 1), gbps=round(nbytes / us / 1e3, 1),
                              frac_of_copy_peak=round(nbytes / us / 1e3 / peak, 3), roofline_us=round(nbytes / peak / 1e3, 1))), flush=True)


if __name__ == '__main__':
    main()
