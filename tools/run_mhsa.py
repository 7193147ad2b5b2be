"""Run the MHSA kernels alone (B, T, H from argv) — the command wrapped by ncu for per-kernel captures."""
import sys
import torch
sys.path.insert(0, '.')
from gaviko_b200 import ops
B, T, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 1033, 12)))
bwd = '--bwd' in sys.argv
torch.manual_seed(0)
qkv = (torch.randn(B * T, 3 * H * 64, device='cuda') * 1.5).bfloat16()
do = torch.randn(B * T, H * 64, device='cuda').bfloat16()
for _ in range(3):
    out, lse = ops.mhsa_fwd(qkv, B, T, H, 0.125)
    if bwd:
        ops.mhsa_bwd(qkv, out, lse, do, B, T, H, 0.125)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out, lse = ops.mhsa_fwd(qkv, B, T, H, 0.125)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print(f'mhsa_fwd B={B} T={T} H={H}: {t*1e3:.1f} us  {4*B*H*T*T*64/t/1e9:.1f} TFLOP/s')
if bwd:
    e0.record()
    for _ in range(10):
        ops.mhsa_bwd(qkv, out, lse, do, B, T, H, 0.125)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    print(f'mhsa_bwd: {t*1e3:.1f} us  {10*B*H*T*T*64/t/1e9:.1f} TFLOP/s (algorithmic 2.5x fwd)')
