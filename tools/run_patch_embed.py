"""Debug / demo driver of gvk_patch_embed: one-hot weights expose which volume element lands in which (token, k) slot."""
import sys, torch
sys.path.insert(0,'.')
from gaviko_b200 import ops
B,C,D,H,W,fp,dim,T_extra = [int(x) for x in (sys.argv[1:9] if len(sys.argv) > 8 else (1,1,12,32,32,12,128,0))]
koff = int(sys.argv[9]) if len(sys.argv) > 9 else 0
ps=16
torch.manual_seed(0)
img = torch.arange(B*C*D*H*W, device='cuda', dtype=torch.float32).reshape(B,C,D,H,W) % 8192      # exactly representable in tf32
K = C*fp*ps*ps
w = torch.zeros(dim, K, device='cuda')
for n in range(dim):
    w[n, (koff + n) % K] = 1.0
n_tok = (D // fp) * (H // ps) * (W // ps)
T = n_tok + T_extra
g = torch.full((B * T, dim), 7.0, device='cuda')
print(ops.patch_embed(img, fp, ps, w, None, None, g, T, T_extra))
torch.cuda.synchronize()
patches = ops.patch_gather(img, fp, ps, torch.float32)            # [B*n_tok, K]
ref = patches[:, [(koff + n) % K for n in range(dim)]]
got = g.view(B, T, dim)[:, T_extra:].reshape(B * n_tok, dim)
bad = (got != ref)
print('mismatches', int(bad.sum()), 'of', bad.numel())
if bad.any():
    idx = bad.nonzero()[:12]
    for t, n in idx.tolist():
        print(f'  tok {t} n {n}: got {got[t, n].item():.0f} expected {ref[t, n].item():.0f}')
    print('rows with errors:', bad.any(1).nonzero().flatten()[:40].tolist())
    print('cols with errors:', bad.any(0).nonzero().flatten()[:64].tolist())
if '--dump' in sys.argv:
    # where did each output come from?  img values are unique (mod 8192): locate got[t, n] in the patch matrix
    flat = patches[:n_tok].reshape(-1)
    for t in range(min(n_tok, 12)):
        row = []
        for n in (0, 1, 15, 16, 17, 31):
            v = got[t, n].item()
            hit = (patches[:n_tok] == v).nonzero()
            row.append(f'n{n}:{v:.0f}<-' + (f'tok{hit[0][0].item()}/k{hit[0][1].item()}' if len(hit) else 'none'))
        print(f'tok {t}: ' + '  '.join(row))
