"""Run the hot-path GEMM shapes alone (the command wrapped by ncu for per-kernel captures): fc2 forward, fc2 dgrad (MUL_AUX), qkv forward."""
import sys
import torch
sys.path.insert(0, '.')
from gaviko_b200 import ops
B, T = 32, 1033
M = B * T
torch.manual_seed(0)
x768 = (torch.randn(M, 768, device='cuda')).bfloat16()
x3072 = (torch.randn(M, 3072, device='cuda')).bfloat16()
w1 = (torch.randn(3072, 768, device='cuda') * 0.03).bfloat16()
w2 = (torch.randn(768, 3072, device='cuda') * 0.02).bfloat16()
wq = (torch.randn(2304, 768, device='cuda') * 0.03).bfloat16()
wo = (torch.randn(768, 768, device='cuda') * 0.03).bfloat16()
b1 = torch.randn(3072, device='cuda'); b2 = torch.randn(768, device='cuda')
res = torch.randn(M, 768, device='cuda')
aux = torch.empty(M, 3072, device='cuda', dtype=torch.bfloat16)
cases = {
    'fc1 fwd (bias+GELU, saves gelu\')': (lambda: ops.gemm(x768, w1, bias=b1, act=ops.ACT_GELU_SAVE_GRAD, aux=aux, out_dtype=torch.bfloat16), 2.0 * M * 768 * 3072),
    'fc2 fwd (bias+residual, fp32 out)': (lambda: ops.gemm(x3072, w2, bias=b2, res1=res), 2.0 * M * 768 * 3072),
    'fc2 dgrad (x gelu\')': (lambda: ops.gemm(x768, w1, act=ops.ACT_MUL_AUX, aux=aux, out_dtype=torch.bfloat16), 2.0 * M * 768 * 3072),
    'fc1 shape, plain bf16 store': (lambda: ops.gemm(x768, w1, out_dtype=torch.bfloat16), 2.0 * M * 768 * 3072),
    'fc1 shape, bias+GELU (no save)': (lambda: ops.gemm(x768, w1, bias=b1, act=ops.ACT_GELU, out_dtype=torch.bfloat16), 2.0 * M * 768 * 3072),
    'fc1 dgrad (K=3072, bf16 out)': (lambda: ops.gemm(x3072, w2, out_dtype=torch.bfloat16), 2.0 * M * 768 * 3072),
    'out-proj fwd (bias+residual)': (lambda: ops.gemm(x768, wo, bias=b2, res1=res), 2.0 * M * 768 * 768),
    'qkv fwd (bf16 out)': (lambda: ops.gemm(x768, wq, out_dtype=torch.bfloat16), 2.0 * M * 768 * 2304),
}
for name, (fn, flops) in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    print(f'{name}: {t * 1e3:.1f} us  {flops / t / 1e9:.0f} TFLOP/s')
