"""GPU: tcgen05 flash attention forward/backward (bf16) against an fp64 torch restatement of model/vision_transformer.py:65-71."""
import pytest
import torch

from gaviko_b200 import ops

pytestmark = pytest.mark.gpu


def _ref(qkv, B, T, H, scale):
    D = 64
    q, k, v = qkv.double().view(B, T, 3, H, D).permute(2, 0, 3, 1, 4)
    a = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    return (a @ v).transpose(1, 2).reshape(B * T, H * D), torch.logsumexp(q @ k.transpose(-1, -2) * scale, -1)


@pytest.mark.parametrize('B,T,H', [(2, 333, 3), (1, 1033, 12), (3, 64, 2), (2, 65, 1), (2, 128, 2), (1, 129, 3), (2, 9, 3), (2, 1009, 3)])
def test_mhsa_fwd_bwd(B, T, H):
    torch.manual_seed(B * 1000 + T + H)
    dim = H * 64
    qkv = (torch.randn(B * T, 3 * dim, device='cuda') * 1.5).bfloat16()
    out, lse = ops.mhsa_fwd(qkv, B, T, H, 0.125)
    qr = qkv.double().requires_grad_(True)
    ref, lse_ref = _ref(qr, B, T, H, 0.125)
    err = (out.double() - ref.detach()).abs().max().item()
    assert err < 2e-2, ('fwd', err)
    assert (lse.view(B, H, T).double() - lse_ref.detach()).abs().max().item() < 1e-3
    do = torch.randn(B * T, dim, device='cuda').bfloat16()
    ref.backward(do.double())
    dqkv = ops.mhsa_bwd(qkv, out, lse, do, B, T, H, 0.125)
    g = qr.grad
    for name, sl in (('dq', slice(0, dim)), ('dk', slice(dim, 2 * dim)), ('dv', slice(2 * dim, 3 * dim))):
        a, b = dqkv[:, sl].double(), g[:, sl]
        rel = (a - b).norm().item() / b.norm().item()
        assert rel < 2e-2, (name, rel, (a - b).abs().max().item())


@pytest.mark.parametrize('B,T,H,p', [(2, 333, 3, 0.1), (1, 1033, 2, 0.2), (2, 129, 2, 0.5), (1, 16, 1, 0.1)])
def test_mhsa_dropout_matches_the_documented_mask_fwd_and_bwd(B, T, H, p):
    """Attention-probability dropout on the tcgen05 path (model/vision_transformer.py:50,69): the forward's mask equals the host restatement
    of the Philox rule in gvk.h, and both backward kernels replay exactly that mask (gradients of the masked fp64 reference)."""
    from helpers import mhsa_keep_mask
    torch.manual_seed(7 + T)
    dim, seed = H * 64, 0x1234567887654321 + T
    qkv = (torch.randn(B * T, 3 * dim, device='cuda') * 1.2).bfloat16()
    keep, scale = mhsa_keep_mask(B, H, T, p, seed)
    keep = keep.cuda()
    out, lse = ops.mhsa_fwd(qkv, B, T, H, 0.125, drop_p=p, seed=seed)
    qr = qkv.double().requires_grad_(True)
    q, k, v = qr.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    dots = q @ k.transpose(-1, -2) * 0.125
    a = torch.softmax(dots, -1) * keep * scale
    ref = (a @ v).transpose(1, 2).reshape(B * T, dim)
    assert (lse.view(B, H, T).double() - torch.logsumexp(dots, -1).detach()).abs().max().item() < 1e-3      # lse is that of the un-dropped softmax
    err = (out.double() - ref.detach()).abs().max().item()
    assert err < 3e-2, ('fwd', err)
    do = torch.randn(B * T, dim, device='cuda').bfloat16()
    ref.backward(do.double())
    dqkv = ops.mhsa_bwd(qkv, out, lse, do, B, T, H, 0.125, drop_p=p, seed=seed)
    for name, sl in (('dq', slice(0, dim)), ('dk', slice(dim, 2 * dim)), ('dv', slice(2 * dim, 3 * dim))):
        x, y = dqkv[:, sl].double(), qr.grad[:, sl]
        rel = (x - y).norm().item() / y.norm().item()
        assert rel < 2e-2, (name, rel)
    frac = keep.float().mean().item()
    assert abs(frac - 1.0 / scale) < 4.0 / (B * H * T * T) ** 0.5 + 1e-3, frac


def test_mhsa_dropout_is_unbiased_and_seeded():
    B, T, H, p = 4, 200, 2, 0.25
    torch.manual_seed(3)
    qkv = torch.randn(B * T, 3 * H * 64, device='cuda').bfloat16()
    base, _ = ops.mhsa_fwd(qkv, B, T, H, 0.125)
    acc = torch.zeros_like(base, dtype=torch.float32)
    outs = []
    for s in range(24):
        o, _ = ops.mhsa_fwd(qkv, B, T, H, 0.125, drop_p=p, seed=1000 + s)
        acc += o.float()
        outs.append(o)
    mean = acc / 24
    assert (mean - base.float()).norm().item() / base.float().norm().item() < 0.2        # E[dropout(P) V] = P V
    assert not torch.equal(outs[0], outs[1])
    again, _ = ops.mhsa_fwd(qkv, B, T, H, 0.125, drop_p=p, seed=1000)
    assert torch.equal(again, outs[0])
