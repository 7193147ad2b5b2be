"""Input rescale: oracle known answers on CPU, CUDA kernel bit-exact against the oracle on GPU."""
import numpy as np
import pytest
import torch

from oracle.rescale_oracle import rescale_intensity as oracle_rescale


def test_oracle_known_answers():
    x = np.array([[2.0, 4.0], [6.0, 10.0]], dtype=np.float32)
    np.testing.assert_array_equal(oracle_rescale(x), np.array([[0.0, 0.25], [0.5, 1.0]], dtype=np.float32))
    np.testing.assert_array_equal(oracle_rescale(x, -1.0, 1.0), np.array([[-1.0, -0.5], [0.0, 1.0]], dtype=np.float32))
    c = np.full((3, 4), 7.5, dtype=np.float32)
    np.testing.assert_array_equal(oracle_rescale(c), c)                    # constant image: returned unchanged
    i16 = np.array([-100, 0, 300], dtype=np.int16)
    np.testing.assert_array_equal(oracle_rescale(i16), np.array([0.0, 0.25, 1.0], dtype=np.float32))
    r = np.random.default_rng(0).normal(size=(5, 7, 3)).astype(np.float32) * 37 + 11
    y = oracle_rescale(r)
    assert y.dtype == np.float32 and y.min() == 0.0 and y.max() == 1.0
    # fp32 operation order: (x - min) / range, not x * (1 / range) - min / range
    want = ((r - r.min()) / (r.max() - r.min())).astype(np.float32)
    np.testing.assert_array_equal(y, want)


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(3, 1, 12, 16, 16), (2, 1, 120, 160, 160), (4, 1, 5, 7, 9), (1, 2, 3, 5, 7)])
def test_rescale_matches_oracle_bit_exact(shape):
    from gaviko_b200.data import RescaleIntensity
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g) * 300 + 40
    x[0].mul_(0.001)                                           # volumes with very different ranges
    want = np.stack([oracle_rescale(v.numpy()) for v in x])
    got = RescaleIntensity()(x.cuda())
    assert got.dtype == torch.float32 and torch.equal(got.cpu(), torch.from_numpy(want))
    want2 = np.stack([oracle_rescale(v.numpy(), -1.0, 3.0) for v in x])
    assert torch.equal(RescaleIntensity(out_min_max=(-1, 3))(x.cuda()).cpu(), torch.from_numpy(want2))
    # bf16 output = the fp32 result rounded once
    assert torch.equal(RescaleIntensity(out_dtype=torch.bfloat16)(x.cuda()).cpu(), torch.from_numpy(want).bfloat16())
    # in place, single volume (C, D, H, W), constant volume passthrough
    xi = x.cuda().clone()
    assert RescaleIntensity()(xi, inplace=True).data_ptr() == xi.data_ptr() and torch.equal(xi.cpu(), torch.from_numpy(want))
    assert torch.equal(RescaleIntensity()(x[-1].cuda()).cpu(), torch.from_numpy(want[-1]))
    c = torch.full(shape, 3.25)
    assert torch.equal(RescaleIntensity()(c.cuda()).cpu(), c)


@pytest.mark.gpu
def test_rescale_rejects_what_it_does_not_implement():
    from gaviko_b200._lib import GvkError
    from gaviko_b200.data import RescaleIntensity
    with pytest.raises(NotImplementedError):
        RescaleIntensity(percentiles=(1, 99))
    with pytest.raises(ValueError):
        RescaleIntensity(out_min_max=(1, 0))
    with pytest.raises(GvkError):
        RescaleIntensity()(torch.zeros(2, 1, 4, 4, 4))        # CPU tensor: no fallback
