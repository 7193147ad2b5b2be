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


@pytest.mark.parametrize('name', ['gaviko_t16_small'] + list(VARIANT_CASES) + list(NEXT_CASES))
def test_engine_host_logic_bf16_mode(name, tmp_path, monkeypatch):
    """The bf16-mode branches of the engines (bf16 operands and activations, K-extension operands packed as hi / lo slots, bf16 LayerNorm-output
    gradients, folded SSF operands) through the double: same bars as the GPU bf16 parity tests (logits 2e-2, identical argmax, gradients 2e-2
    globally or the reference's own bf16 deviation where that is larger).  The arithmetic inside each wrapper is the double's fp32, so this checks
    the engines' bookkeeping in that mode, not the kernels."""
    monkeypatch.chdir(tmp_path)
    if name in GAVIKO_CASES:
        kw, batch = GAVIKO_CASES[name]
        with contextlib.redirect_stdout(io.StringIO()):
            from gaviko_b200.model.gaviko import Gaviko
            model = Gaviko(**kw, compute_dtype='bf16')
    else:
        method, kw, batch = dict(VARIANT_CASES, **NEXT_CASES)[name]
        model = build_variant(method, dict(kw, compute_dtype='bf16'))
    golden_fill(model, seed=0)
    model.eval()
    g = load_golden(name)
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'], channels=kw['channels'])
    y = golden_labels(batch, kw['num_classes'])
    with ops_double.install():
        for loss_name, crit in (('focal', FocalLoss(gamma=1.2)), ('ce', CrossEntropyLoss())):
            model.zero_grad(set_to_none=True)
            logits = model(img)
            crit(logits, y).backward()
            assert rel_l2(logits.detach().float(), g['logits']) < 2e-2
            assert logits.argmax(1).tolist() == g['logits'].argmax(1).tolist()
            grads = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
            ref_dev = lambda k: float(g[k]) if k in g else 0.0      # noqa: E731
            tol_g = max(2e-2, 2 * ref_dev(f'refbf16_grad_global_{loss_name}'))
            tol_t = max(0.15, 2 * ref_dev(f'refbf16_grad_worst_{loss_name}'))
            grad_parity(grads, g, loss_name, tol_global=tol_g, tol_tensor=tol_t, floor=1e-2, floor_slack=2.0)


def test_gaviko_bf16_mode_fused_side_passes_host_logic(monkeypatch):
    """bf16 compute mode at dim 768 takes the one-pass forms (gvk_layernorm_fwd_down, gvk_rowproj_up_down, gvk_layernorm_bwd with an output
    projection).  On the double each of them is the composition of the kernels it replaces, so the step with and without them must agree to
    rounding: a wrong layer's weights, a stale d(comb) or a missed in-place update in the engine shows up here without a GPU."""
    kw, batch = GAVIKO_CASES['gaviko_t16_small']
    kw = dict(kw, backbone='vit-b16')                      # dim 768, 12 layers, share_factor 2; 64 + 9 tokens
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'], channels=kw['channels'])
    y = golden_labels(batch, kw['num_classes'])

    def run(fused):
        with contextlib.redirect_stdout(io.StringIO()):
            from gaviko_b200.model.gaviko import Gaviko
            model = Gaviko(**kw, compute_dtype='bf16')
        golden_fill(model, seed=0)
        model.eval()
        calls = []
        with monkeypatch.context() as mp:
            for n in ('layernorm_fwd_down', 'rowproj_up_down', 'layernorm_bwd'):
                f = getattr(ops_double, n)
                mp.setattr(ops_double, n, (lambda f, n: lambda *a, **k: (calls.append((n, 'ow' in k and k['ow'] is not None)), f(*a, **k))[1])(f, n))
            if not fused:
                for n in ('layernorm_fwd_down_supported', 'layernorm_bwd_down_supported', 'rowproj_up_down_supported'):
                    mp.setattr(ops_double, n, lambda *a, **k: False)
            with ops_double.install():
                logits = model(img)
                CrossEntropyLoss()(logits, y).backward()
        return logits.detach(), {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}, calls

    l1, g1, c1 = run(True)
    l0, g0, c0 = run(False)
    depth = 12
    assert c1.count(('layernorm_fwd_down', False)) == depth and c1.count(('rowproj_up_down', False)) == 2 * depth
    assert c1.count(('layernorm_bwd', True)) == depth - 1          # every LayerNorm1 backward but layer 0's projects its output for the next iteration
    assert not any(n in ('layernorm_fwd_down', 'rowproj_up_down') or ow for n, ow in c0)
    assert rel_l2(l1, l0) < 1e-6
    for n in g1:
        assert rel_l2(g1[n], g0[n]) < 1e-5, n


def test_double_is_not_reachable_without_install():
    """The double only exists inside `install()`: outside it a CPU tensor still raises (no CPU fallback)."""
    from gaviko_b200._lib import GvkError
    from gaviko_b200.model.gaviko import Gaviko
    kw, _ = GAVIKO_CASES['gaviko_t16_small']
    with contextlib.redirect_stdout(io.StringIO()):
        m = Gaviko(**kw)
    with pytest.raises(GvkError):
        m(torch.zeros(1, 1, 48, 64, 64))
