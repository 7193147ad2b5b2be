"""GPU: gvk_gemm (tcgen05 bf16 path and exact fp32 path) against torch matmul, through the C ABI."""
import pytest
import torch

from gaviko_b200 import ops

pytestmark = pytest.mark.gpu


def _ref(a, b, bias=None, act=None, res1=None, res2=None, pre=None):
    v = a.double() @ b.double().t()
    if bias is not None:
        v = v + bias.double()
    if act == 'gelu':
        v = torch.nn.functional.gelu(v)
    if act == 'gelu_bwd':
        x = pre.double()
        cdf = 0.5 * (1 + torch.erf(x / 2 ** 0.5))
        pdf = torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5
        v = v * (cdf + x * pdf)
    if res1 is not None:
        v = v + res1.double()
    if res2 is not None:
        v = v + res2.double()
    return v


SHAPES = [(128, 256, 64), (300, 768, 768), (2066, 2304, 768), (1033, 192, 192), (2000, 576, 192), (515, 768, 3072), (97, 3072, 768), (4132, 1024, 4096)]


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float32])
def test_gemm_plain(M, N, K, dt):
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device='cuda').to(dt)
    b = (torch.randn(N, K, device='cuda') / K ** 0.5).to(dt)
    out = ops.gemm(a, b, out_dtype=torch.float32)
    ref = _ref(a, b)
    err = (out.double() - ref).abs().max().item()
    tol = 2e-3 if dt == torch.bfloat16 else 2e-5   # bf16 products are exact in fp32; error is accumulation order only
    assert err < tol * max(1.0, ref.abs().max().item()), (err, ref.abs().max().item())


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float32])
def test_gemm_epilogues(dt):
    torch.manual_seed(0)
    M, N, K = 1033 * 2, 768, 192
    a = torch.randn(M, K, device='cuda').to(dt)
    b = (torch.randn(N, K, device='cuda') / K ** 0.5).to(dt)
    bias = torch.randn(N, device='cuda')
    res1 = torch.randn(M, N, device='cuda')
    res2 = torch.randn(M, N, device='cuda')
    # bias + gelu with saved pre-activation, bf16 output
    aux = torch.empty(M, N, device='cuda', dtype=torch.float32)
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, aux=aux, out_dtype=torch.bfloat16)
    assert (aux.double() - _ref(a, b, bias)).abs().max().item() < 1e-3
    assert (out.double() - _ref(a, b, bias, 'gelu')).abs().max().item() < 3e-2      # bf16 output rounding
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, out_dtype=torch.float32)
    assert (out.double() - _ref(a, b, bias, 'gelu')).abs().max().item() < 1e-3
    # gelu backward epilogue + two residuals, fp32 output
    out = ops.gemm(a, b, act=ops.ACT_GELU_BWD, aux=aux, res1=res1, res2=res2, out_dtype=torch.float32)
    assert (out.double() - _ref(a, b, None, 'gelu_bwd', res1, res2, pre=aux)).abs().max().item() < 1e-3
    # GELU with the derivative saved instead of the pre-activation, and its one-multiply backward epilogue (same result as GELU_BWD)
    pre = _ref(a, b, bias)
    dgelu = 0.5 * (1 + torch.erf(pre / 2 ** 0.5)) + pre * torch.exp(-0.5 * pre * pre) / (2 * torch.pi) ** 0.5
    aux2 = torch.empty(M, N, device='cuda', dtype=dt)
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU_SAVE_GRAD, aux=aux2, out_dtype=dt)
    assert (out.double() - _ref(a, b, bias, 'gelu')).abs().max().item() < (3e-2 if dt == torch.bfloat16 else 1e-3)
    assert (aux2.double() - dgelu).abs().max().item() < (1e-2 if dt == torch.bfloat16 else 1e-3)
    out = ops.gemm(a, b, act=ops.ACT_MUL_AUX, aux=aux2, out_dtype=dt)
    want = _ref(a, b) * aux2.double()
    assert (out.double() - want).abs().max().item() < (3e-2 if dt == torch.bfloat16 else 1e-3) * max(1.0, want.abs().max().item())
    # in-place residual update (out aliases res1), as the residual stream does
    r = res1.clone()
    ops.gemm(a, b, bias=bias, res1=r, out=r)
    assert (r.double() - _ref(a, b, bias, None, res1)).abs().max().item() < 1e-3


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float32])
def test_gemm_patch_embed_epilogue(dt):
    """Row remap + positional add + dual store: the a1+a2 fusion (model/gaviko.py:532-548)."""
    torch.manual_seed(1)
    B, Np, P, K, N = 3, 64, 8, 3072, 192
    T = P + 1 + Np
    a = torch.rand(B * Np, K, device='cuda').to(dt)
    b = (torch.randn(N, K, device='cuda') / K ** 0.5).to(dt)
    bias = torch.randn(N, device='cuda')
    pos = torch.randn(Np, N, device='cuda')
    g = torch.zeros(B * T, N, device='cuda')
    loc = torch.empty(B * Np, N, device='cuda')
    ops.gemm(a, b, bias=bias, pos=pos, rows_per_batch=Np, out_batch_rows=T, out_row_offset=P + 1, out=g, out2=loc)
    ref = (_ref(a, b, bias).view(B, Np, N) + pos.double()).float()
    assert (loc.view(B, Np, N) - ref).abs().max().item() < 1e-3
    assert (g.view(B, T, N)[:, P + 1:] - ref).abs().max().item() < 1e-3
    assert g.view(B, T, N)[:, :P + 1].abs().max().item() == 0


def test_gemm_rejects_cpu_tensors():
    from gaviko_b200._lib import GvkError
    with pytest.raises(GvkError):
        ops.gemm(torch.zeros(4, 64), torch.zeros(4, 64))


def test_k_extension_with_split_operands():
    """The rank-r fp32 product rides on a bf16 GEMM as one extra K block: [A | (hi, lo, hi)(c)] [B | (hi, hi, lo)(wu)]^T = A B^T + c wu^T with the
    rank-r term accurate to ~2^-16 (three bf16 products), far inside the tf32 accuracy of the separate up-projection kernel it replaces."""
    torch.manual_seed(5)
    M, N, K, r = 1033, 256, 128, 20
    c = torch.randn(M, r, device='cuda')
    wu = torch.randn(N, r, device='cuda') / r ** 0.5
    a = torch.zeros(M, K + 64, device='cuda', dtype=torch.bfloat16)       # A part zero: isolates the rank-r term
    b = torch.zeros(N, K + 64, device='cuda', dtype=torch.bfloat16)
    ops.split_pack_bf16(c, a[:, K:], 0b010)
    ops.split_pack_bf16(wu, b[:, K:], 0b100)
    hi = c.bfloat16()
    assert torch.equal(a[:, K:K + r], hi) and torch.equal(a[:, K + 2 * r:K + 3 * r], hi) and torch.equal(a[:, K + r:K + 2 * r], (c - hi.float()).bfloat16())
    assert (a[:, K + 3 * r:] == 0).all() and torch.equal(b[:, K + 2 * r:K + 3 * r], (wu - wu.bfloat16().float()).bfloat16())
    out = ops.gemm(a, b)
    want = c.double() @ wu.double().t()
    rel = ((out.double() - want).norm() / want.norm()).item()
    assert rel < 5e-5, rel
    plain = (c.bfloat16().double() @ wu.bfloat16().double().t())          # what a single bf16 product would give
    assert rel < 0.05 * ((plain - want).norm() / want.norm()).item()


@pytest.mark.parametrize('B,C,D,H,W,fp,dim,T_extra', [(2, 1, 120, 160, 160, 12, 768, 33), (3, 1, 24, 64, 48, 12, 192, 9), (1, 2, 12, 32, 32, 6, 128, 0), (5, 1, 36, 160, 160, 12, 256, 1)])
def test_patch_embed_fused_tma_gather(B, C, D, H, W, fp, dim, T_extra):
    """gvk_patch_embed (TMA patch gather -> tf32 tcgen05 GEMM -> bias + positional embedding -> both token streams) against Conv3d in fp64
    (model/gaviko.py:383-385,532-548); tolerance = tf32 operand rounding (2^-11 relative per product, K up to 3072)."""
    from gaviko_b200 import ops
    torch.manual_seed(B * 7 + dim)
    ps = 16
    img = torch.rand(B, C, D, H, W, device='cuda') * 2 - 0.7
    w = torch.randn(dim, C, fp, ps, ps, device='cuda') * 0.02
    bias = torch.randn(dim, device='cuda')
    n_tok = (D // fp) * (H // ps) * (W // ps)
    pos = torch.randn(n_tok, dim, device='cuda')
    T = n_tok + T_extra
    g = torch.full((B * T, dim), 7.0, device='cuda')
    loc = torch.empty((B * n_tok, dim), device='cuda')
    assert ops.patch_embed(img, fp, ps, w.reshape(dim, -1).contiguous(), bias, pos, g, T, T_extra, out2=loc)
    ref = torch.nn.functional.conv3d(img.double(), w.double(), bias.double(), stride=(fp, ps, ps)).flatten(2).transpose(1, 2) + pos.double()
    got = g.view(B, T, dim)[:, T_extra:].double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, err
    assert torch.equal(loc.view(B, n_tok, dim), g.view(B, T, dim)[:, T_extra:])
    if T_extra:
        assert bool((g.view(B, T, dim)[:, :T_extra] == 7.0).all())        # the prompt / cls rows are not touched
