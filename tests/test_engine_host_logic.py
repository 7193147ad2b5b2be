"""CPU: the HOST logic of the engines (call order, in-place conventions, token bookkeeping, gradient routing) against the golden vectors of the
live reference, with every C-ABI wrapper replaced by its torch restatement (tests/ops_double.py — a test double, never shipped).  The GPU tests
(-m gpu) run the same cases through the real kernels; this file makes an engine bug visible without a GPU."""
import contextlib
import io

import pytest
import torch

import ops_double
from gaviko_b200.losses.focal_loss import CrossEntropyLoss, FocalLoss
from oracle.cases import GAVIKO_CASES, NEXT_CASES, VARIANT_CASES
from oracle.golden_fill import golden_fill, golden_labels, golden_volume

from helpers import grad_parity, load_golden, rel_l2
from variant_factory import build_variant


def _check(model, kw, batch, g, tol_g=5e-5):
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'], channels=kw['channels'])
    y = golden_labels(batch, kw['num_classes'])
    with ops_double.install():
        for loss_name, crit in (('focal', FocalLoss(gamma=1.2)), ('ce', CrossEntropyLoss())):
            model.zero_grad(set_to_none=True)
            logits = model(img)
            loss = crit(logits, y)
            loss.backward()
            assert rel_l2(logits.detach(), g['logits']) < 2e-5
            assert abs(loss.item() - float(g[f'loss_{loss_name}'])) < 1e-5
            grads = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
            grad_parity(grads, g, loss_name, tol_global=tol_g, tol_tensor=1e-3)


def test_gaviko_engine_host_logic():
    kw, batch = GAVIKO_CASES['gaviko_t16_small']
    with contextlib.redirect_stdout(io.StringIO()):
        from gaviko_b200.model.gaviko import Gaviko
        model = Gaviko(**kw, compute_dtype='fp32')
    golden_fill(model, seed=0)
    model.eval()
    _check(model, kw, batch, load_golden('gaviko_t16_small'))


@pytest.mark.parametrize('name', list(VARIANT_CASES) + list(NEXT_CASES))
def test_variant_engine_host_logic(name, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    method, kw, batch = dict(VARIANT_CASES, **NEXT_CASES)[name]
    model = build_variant(method, dict(kw, compute_dtype='fp32'))
    golden_fill(model, seed=0)
    model.eval()
    _check(model, kw, batch, load_golden(name))


def test_double_is_not_reachable_without_install():
    """The double only exists inside `install()`: outside it a CPU tensor still raises (no CPU fallback)."""
    from gaviko_b200._lib import GvkError
    from gaviko_b200.model.gaviko import Gaviko
    kw, _ = GAVIKO_CASES['gaviko_t16_small']
    with contextlib.redirect_stdout(io.StringIO()):
        m = Gaviko(**kw)
    with pytest.raises(GvkError):
        m(torch.zeros(1, 1, 48, 64, 64))
