import torch
M=33056
for (K,N,name) in ((3072,768,'fc2 fwd'),(768,3072,'fc1 fwd'),(768,2304,'qkv')):
    a=torch.randn(M,K,device='cuda',dtype=torch.bfloat16); w=torch.randn(N,K,device='cuda',dtype=torch.bfloat16)
    for _ in range(3): torch.matmul(a,w.t())
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch.matmul(a,w.t())
    e1.record(); torch.cuda.synchronize()
    t=e0.elapsed_time(e1)/20
    print(f'cuBLAS {name} M={M} K={K} N={N}: {t*1e3:.1f} us {2*M*K*N/t/1e9:.0f} TFLOP/s (plain bf16 out, no epilogue)')
