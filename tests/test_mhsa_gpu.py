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
