"""Bring-up probe for the tcgen05 attention kernels: error report per stage (and timing)."""
import sys, os
import torch
sys.path.insert(0, '.')
from gaviko_b200 import ops

def ref(qkv, B, T, H, scale):
    q, k, v = qkv.double().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) * scale
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, H * 64)

print('GVK_FA_DESC =', os.environ.get('GVK_FA_DESC'))
torch.manual_seed(0)
for (B, T, H) in [(1, 64, 1), (1, 128, 1), (2, 333, 3), (1, 1033, 12)]:
    dim = H * 64
    qkv = (torch.randn(B * T, 3 * dim, device='cuda')).bfloat16()
    out, lse = ops.mhsa_fwd(qkv, B, T, H, 0.125)
    torch.cuda.synchronize()
    qr = qkv.double().requires_grad_(True)
    r = ref(qr, B, T, H, 0.125)
    print(f'B{B} T{T} H{H}: fwd max err {(out.double() - r.detach()).abs().max().item():.3e}', flush=True)
    do = torch.randn(B * T, dim, device='cuda').bfloat16()
    r.backward(do.double())
    dqkv = ops.mhsa_bwd(qkv, out, lse, do, B, T, H, 0.125)
    torch.cuda.synchronize()
    for name, sl in (('dq', slice(0, dim)), ('dk', slice(dim, 2 * dim)), ('dv', slice(2 * dim, 3 * dim))):
        a, b = dqkv[:, sl].double(), qr.grad[:, sl]
        print(f'    {name} rel {((a - b).norm() / b.norm()).item():.3e}', flush=True)
if len(sys.argv) > 1 and sys.argv[1] == 'time':
    B, T, H = 32, 1033, 12
    qkv = torch.randn(B * T, 3 * H * 64, device='cuda').bfloat16(); do = torch.randn(B * T, H * 64, device='cuda').bfloat16()
    out, lse = ops.mhsa_fwd(qkv, B, T, H, 0.125)
    for fn, name, fl in ((lambda: ops.mhsa_fwd(qkv, B, T, H, 0.125), 'fwd', 4), (lambda: ops.mhsa_bwd(qkv, out, lse, do, B, T, H, 0.125), 'bwd', 10)):
        for _ in range(2): fn()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'{name}: {ms:.3f} ms  = {fl * B * H * T * T * 64 / ms / 1e9:.0f} TFLOP/s (algorithmic {fl}*B*H*T^2*64)')
