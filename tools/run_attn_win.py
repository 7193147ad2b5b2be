"""Time the window attention core (LocalSelfAttention, model/gaviko.py:229-244) in its exact-SIMT and tf32 tensor-core forms.
python tools/run_attn_win.py [--batch 64] [--drop 0.2]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaviko_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--drop', type=float, default=0.2)
    ap.add_argument('--iters', type=int, default=20)
    a = ap.parse_args()
    B, N, r, dhw, k = a.batch, 1000, 20, (10, 10, 10), (6, 6, 6)
    torch.manual_seed(0)
    qkv = torch.randn(B * N, 3 * r, device='cuda')
    do = torch.randn(B * N, r, device='cuda')
    for name, prec in (('simt fp32', ops.PREC_FP32), ('mma tf32', ops.PREC_TF32)):
        kw = dict(q_off=0, k_off=r, v_off=2 * r, scale=768 ** -0.5, window=k, grid=dhw, drop_p=a.drop, seed=3, prec=prec)
        o, lse = ops.attn_simt_fwd(qkv, B, N, 1, r, **kw)
        ops.attn_simt_bwd(qkv, o, lse, do, B, N, 1, r, **kw)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(a.iters):
            o, lse = ops.attn_simt_fwd(qkv, B, N, 1, r, **kw)
        ev[1].record()
        for _ in range(a.iters):
            ops.attn_simt_bwd(qkv, o, lse, do, B, N, 1, r, **kw)
        ev[2].record()
        torch.cuda.synchronize()
        print(f'{name}: B={B} drop={a.drop}  fwd {ev[0].elapsed_time(ev[1]) / a.iters * 1e3:.1f} us   bwd {ev[1].elapsed_time(ev[2]) / a.iters * 1e3:.1f} us')


if __name__ == '__main__':
    main()
