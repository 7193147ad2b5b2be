"""Diagnostic run for the first GPU bring-up of the tcgen05 GEMM: prints errors and timings instead of asserting."""
import sys, time
import torch
sys.path.insert(0, '.')
from gaviko_b200 import ops

torch.manual_seed(0)
dev = 'cuda'
print(torch.cuda.get_device_name(0))
for dt in (torch.float32, torch.bfloat16):
    for (M, N, K) in [(128, 128, 64), (128, 256, 64), (128, 256, 128), (256, 256, 768), (300, 768, 768), (2066, 2304, 768), (1033, 192, 192)]:
        a = torch.randn(M, K, device=dev).to(dt)
        b = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
        try:
            out = ops.gemm(a, b, out_dtype=torch.float32)
            torch.cuda.synchronize()
            ref = a.double() @ b.double().t()
            err = (out.double() - ref).abs().max().item()
            print(f'{dt} M{M} N{N} K{K}: max err {err:.3e} (ref max {ref.abs().max().item():.2f})', flush=True)
            if err > 1e-2 and M <= 256:
                bad = ((out.double() - ref).abs() > 1e-2)
                rows = bad.any(1).nonzero().flatten()[:16].tolist(); cols = bad.any(0).nonzero().flatten()[:16].tolist()
                print('   bad rows', rows, 'bad cols', cols, 'frac', bad.float().mean().item())
                print('   out[0,:8]', out[0, :8].tolist()); print('   ref[0,:8]', ref[0, :8].tolist())
        except Exception as ex:
            print(f'{dt} M{M} N{N} K{K}: EXC {ex}', flush=True)

# timing at benchmark shapes
def bench(M, N, K, **kw):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, b, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): ops.gemm(a, b, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    e0.record()
    for _ in range(10): torch.matmul(a, b.t(), out=out)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print(f'bench M{M} N{N} K{K}: gvk {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TFLOP/s | torch.matmul {ms2:.3f} ms = {2*M*N*K/ms2/1e9:.0f} TFLOP/s', flush=True)

try:
    for shp in [(33056, 2304, 768), (33056, 768, 768), (33056, 3072, 768), (33056, 768, 3072)]:
        bench(*shp)
    M = 33056
    bias = torch.randn(3072, device=dev)
    bench(M, 3072, 768, bias=bias, act=ops.ACT_GELU)
except Exception as ex:
    print('bench EXC', ex)
