"""Time the rank-r row kernels on the hot-path shapes (M = batch * 1033 rows, dim 768, r 20) against the HBM roofline.
python tools/run_rowops.py [--batch 64] [--only wgrad]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaviko_b200 import ops  # noqa: E402


def timeit(fn, iters, flush, mode):
    fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    for i in range(iters):
        if mode == 'dirty':
            flush.add_(1.0)                  # > L2 and leaves L2 full of dirty lines the timed kernel has to evict (as inside a real step)
        elif mode == 'clean':
            flush.sum()                      # > L2, clean lines only
        torch.cuda._sleep(400000)            # keep the GPU busy while the host enqueues: the events then bracket device time only
        ev[2 * i].record()
        fn()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(iters))
    return ts[len(ts) // 2] * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--only', default='')
    ap.add_argument('--flush', default='dirty', choices=['dirty', 'clean', 'none'])
    a = ap.parse_args()
    M, dim, r = a.batch * 1033, 768, 20
    peak = 6550.0
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    dev = 'cuda'
    torch.manual_seed(0)
    x = torch.randn(M, dim, device=dev)
    dy = torch.randn(M, dim, device=dev)
    dyb = dy.bfloat16()
    c = torch.randn(M, r, device=dev)
    w = torch.randn(r, dim, device=dev) * 0.05
    wu = torch.randn(dim, r, device=dev) * 0.05
    gamma, beta = torch.rand(dim, device=dev) + 0.5, torch.randn(dim, device=dev)
    mean, rstd = x.mean(1), 1.0 / x.std(1)
    dw = torch.zeros(dim, r, device=dev)
    w2 = torch.randn(60, r, device=dev)
    out = torch.empty_like(x)
    out_lp = torch.empty(M, dim, device=dev, dtype=torch.bfloat16)
    flush = torch.zeros(64 << 20, device=dev)   # 256 MB > L2
    P = ops.PREC_TF32
    f32 = M * dim * 4
    cases = {
        'down (LN, w2)': (lambda: ops.rowproj_down(x, w, None, ln=(gamma, beta), w2=w2, prec=P), f32),
        'down (plain)': (lambda: ops.rowproj_down(x, w, None, prec=P), f32),
        'down (transposed)': (lambda: ops.rowproj_down(x, wu, transposed=True, prec=P), f32),
        'up (res)': (lambda: ops.rowproj_up(c, wu, None, res=x, out=out, prec=P), 2 * f32),
        'up (res, bf16 copy)': (lambda: ops.rowproj_up(c, wu, None, res=x, out=out, out_lp=out_lp, prec=P), 2.5 * f32),
        'up (no res)': (lambda: ops.rowproj_up(c, wu, None, out=out, prec=P), f32),
        'wgrad': (lambda: ops.skinny_wgrad(c, x, dw=dw, dw_layout='dr', prec=P), f32),
        'wgrad (LN)': (lambda: ops.skinny_wgrad(c, x, dw=dw, dw_layout='dr', ln=(gamma, beta, mean, rstd), prec=P), f32),
        'ln_bwd (dy fp32 + dres)': (lambda: ops.layernorm_bwd(x, gamma, mean, rstd, dy=dy, dres=dy, dx=out), 4 * f32),
        'ln_bwd (dz rank-r + dres)': (lambda: ops.layernorm_bwd(x, gamma, mean, rstd, dz=c, w=w, dres=dy, dx=out), 3 * f32),
        'ln_fwd (bf16 out)': (lambda: ops.layernorm_fwd(x, gamma, beta, out_dtype=torch.bfloat16), 1.5 * f32),
    }
    for name, (fn, nbytes) in cases.items():
        if a.only and a.only not in name:
            continue
        try:
            us = timeit(fn, a.iters, flush, a.flush)
        except Exception as e:      # noqa: BLE001
            print(f'{name:28s} failed: {e}')
            continue
        gbs = nbytes / us / 1e3
        print(f'{name:28s} {us:8.1f} us   {gbs:7.0f} GB/s algorithmic   {gbs / peak * 100:5.1f} % of {peak:.0f}')


if __name__ == '__main__':
    main()
