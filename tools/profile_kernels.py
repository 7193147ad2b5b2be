"""Every hot kernel of the training step once, on the step's own shapes (B volumes of ViT-B/16 GAViKO) — the command wrapped by
`ncu --set full` for the per-kernel captures under profiles/ (each op runs twice; the summary keeps the last launch of each kernel)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaviko_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
a = ap.parse_args()
B, T, N, H, dim, mlp, r = a.batch, 1033, 1000, 12, 768, 3072, 20
M, Ml = B * T, B * N
dev = 'cuda'
torch.manual_seed(0)
bf = torch.bfloat16
x768 = torch.randn(M, dim, device=dev).to(bf)
x3072 = torch.randn(M, mlp, device=dev).to(bf)
w1 = (torch.randn(mlp, dim, device=dev) * 0.03).to(bf)
w2 = (torch.randn(dim, mlp, device=dev) * 0.02).to(bf)
wq = (torch.randn(3 * dim, dim, device=dev) * 0.03).to(bf)
b1, b2 = torch.randn(mlp, device=dev), torch.randn(dim, device=dev)
res = torch.randn(M, dim, device=dev)
aux = torch.empty(M, mlp, device=dev, dtype=bf)
qkv = (torch.randn(M, 3 * dim, device=dev) * 1.5).to(bf)
do = torch.randn(M, dim, device=dev).to(bf)
loc = torch.randn(Ml, dim, device=dev)
c = torch.randn(M, r, device=dev)
wd = torch.randn(r, dim, device=dev) * 0.05
wu = torch.randn(dim, r, device=dev) * 0.05
wqkv_l = torch.randn(3 * r, r, device=dev)
gamma, beta = torch.rand(dim, device=dev) + 0.5, torch.randn(dim, device=dev)
mean, rstd = res.mean(1), 1.0 / res.std(1)
dw = torch.zeros(dim, r, device=dev)
out = torch.empty_like(res)
out_lp = torch.empty(M, dim, device=dev, dtype=bf)
dy32 = torch.randn(M, dim, device=dev)
qkv_l = torch.randn(Ml, 3 * r, device=dev)
do_l = torch.randn(Ml, r, device=dev)
P = ops.PREC_TF32
win = dict(q_off=0, k_off=r, v_off=2 * r, scale=dim ** -0.5, window=(6, 6, 6), grid=(10, 10, 10), drop_p=0.2, seed=3, prec=P)


def step():
    ops.gemm(x768, w1, bias=b1, act=ops.ACT_GELU_SAVE_GRAD, aux=aux, out_dtype=bf)           # fc1 forward
    ops.gemm(x768, w1, act=ops.ACT_MUL_AUX, aux=aux, out_dtype=bf)                             # fc2 dgrad
    ops.gemm(x3072, w2, bias=b2, res1=res)                                                     # fc2 forward
    ops.gemm(x768, wq, out_dtype=bf)                                                           # qkv forward
    o, lse = ops.mhsa_fwd(qkv, B, T, H, 0.125)
    ops.mhsa_bwd(qkv, o, lse, do, B, T, H, 0.125)
    ol, lsel = ops.attn_simt_fwd(qkv_l, B, N, 1, r, **win)
    ops.attn_simt_bwd(qkv_l, ol, lsel, do_l, B, N, 1, r, **win)
    ops.rowproj_down(loc, wd, None, ln=(gamma, beta), w2=wqkv_l, prec=P)
    ops.rowproj_up(c, wu, None, res=res, out=out, out_lp=out_lp, prec=P)
    ops.skinny_wgrad(c, res, dw=dw, dw_layout='dr', prec=P)
    ops.layernorm_bwd(res, gamma, mean, rstd, dy=dy32, dres=out, dx=dy32, dx_lp=out_lp)
    ops.layernorm_fwd(res, gamma, beta, out_dtype=bf)


step()
torch.cuda.synchronize()
step()
torch.cuda.synchronize()
print('ok')
