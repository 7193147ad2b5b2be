"""Timeline of CTA 0 of the pipelined attention-backward dQ kernel (GVK_PIPE_DBG=4): clock deltas per tag for the issuer and the two softmax groups.
    GVK_PIPE_DBG=4 python tools/mhsa_trace.py [B T H]"""
import ctypes as C, os, sys
import numpy as np
import torch
sys.path.insert(0, '.')
os.environ.setdefault('GVK_PIPE_DBG', '4')      # 4: dQ kernel, 16: dK/dV kernel, 32: forward kernel (roles 1, 2 = softmax groups of tile A, B)
from gaviko_b200 import ops, _lib as L
B, T, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 1033, 12)))
qkv = (torch.randn(B * T, 3 * H * 64, device='cuda') * 1.5).bfloat16()
do = torch.randn(B * T, H * 64, device='cuda').bfloat16()
out, lse = ops.mhsa_fwd(qkv, B, T, H, 0.125)
for _ in range(2):
    if int(os.environ['GVK_PIPE_DBG']) & 32:
        ops.mhsa_fwd(qkv, B, T, H, 0.125)
    else:
        ops.mhsa_bwd(qkv, out, lse, do, B, T, H, 0.125)
torch.cuda.synchronize()
N = 2048
buf = np.zeros((4, N, 2), dtype=np.uint32)
fn = L.lib().gvk_debug_trace
fn(buf.ctypes.data_as(C.POINTER(C.c_uint32)), buf.size)
names = ['score issuer', 'softmax grp0', 'softmax grp1', 'accumulator issuer']
t00 = min(int(buf[r, 0, 1]) for r in range(4) if buf[r, 0, 0])
for role in (0, 3, 1, 2):
    ev = buf[role]
    n = int((ev[:, 0] != 0).sum())
    print(f'== {names[role]}: {n} events')
    prev = None
    for i in range(min(n, int(os.environ.get('TRACE_ROWS', '80')))):
        tag, clk = int(ev[i, 0]), int(ev[i, 1])
        d = (clk - prev) & 0xffffffff if prev is not None else 0
        print(f'  {tag:#06x}  t={(clk - t00) & 0xffffffff:8d}  +{d}')
        prev = clk
