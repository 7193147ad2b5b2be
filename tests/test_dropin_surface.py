"""CPU: the drop-in boundary — parameter names / shapes / trainable set / state_dict keys equal the reference's (recorded in
tests/golden by oracle/make_golden.py); the C-ABI library loads and exports every symbol include/gvk.h declares; struct mirrors
match sizeof() on the C side.  When /root/reference is present (build container) also: seeded construction reproduces the
reference's initial weights exactly."""
import ctypes

import pytest
import torch

from gaviko_b200 import _lib as L
from oracle import refload
from oracle.cases import GAVIKO_CASES

from helpers import load_golden


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    assert len(L.FUNCTIONS) >= 20
    for fn in L.FUNCTIONS + ['gvk_last_error', 'gvk_launch_count', 'gvk_struct_size']:
        assert hasattr(lib, fn), fn
    assert lib.gvk_version() >= 100


def test_struct_mirrors_match_c_sizeof():
    lib = L.lib()
    lib.gvk_struct_size.restype = ctypes.c_longlong
    for name, st in L.STRUCTS.items():
        assert lib.gvk_struct_size(name.encode()) == ctypes.sizeof(st), name


@pytest.mark.parametrize('name', list(GAVIKO_CASES))
def test_gaviko_names_shapes_match_reference(name):
    from gaviko_b200.model.gaviko import Gaviko
    kw, _ = GAVIKO_CASES[name]
    g = load_golden(name)
    m = Gaviko(**kw)
    assert [n for n, _ in m.named_parameters()] == g['all_names'].tolist()
    assert [str(tuple(p.shape)) for _, p in m.named_parameters()] == g['all_shapes'].tolist()
    assert [n for n, p in m.named_parameters() if p.requires_grad] == g['trainable_names'].tolist()
    assert list(m.state_dict().keys()) == g['state_dict_keys'].tolist()
    assert m.train() is None and m.training is True
    m.eval()
    assert m.training is True and not m.transformer.training      # reference quirk, model/gaviko.py:525-528


def test_no_cpu_fallback():
    from gaviko_b200.model.gaviko import Gaviko
    kw, _ = GAVIKO_CASES['gaviko_t16_small']
    with pytest.raises(L.GvkError):
        Gaviko(**kw)(torch.zeros(1, 1, 48, 64, 64))


@pytest.mark.skipif(not refload.available(), reason='live reference only exists in the build container')
def test_seeded_construction_equals_reference_init():
    from gaviko_b200.model.gaviko import Gaviko
    ref = refload.load()
    kw, _ = GAVIKO_CASES['gaviko_t16_small']
    torch.manual_seed(7)
    a = ref.Gaviko(**kw)
    torch.manual_seed(7)
    b = Gaviko(**kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


def _all_variant_cases():
    from oracle.cases import NEXT_CASES, VARIANT_CASES
    return dict(VARIANT_CASES, **NEXT_CASES)


@pytest.mark.parametrize('name', list(_all_variant_cases()))
def test_variant_names_shapes_match_reference(name, tmp_path, monkeypatch):
    """Every --method variant (incl. dvpt) exposes the reference's parameter names, shapes and trainable set (= the checkpoint layout,
    train.py:161-167)."""
    from variant_factory import build_variant
    monkeypatch.chdir(tmp_path)
    method, kw, _ = _all_variant_cases()[name]
    model = build_variant(method, kw)
    g = load_golden(name)
    assert [n for n, _ in model.named_parameters()] == g['all_names'].tolist()
    assert [str(tuple(p.shape)) for _, p in model.named_parameters()] == [s.replace(' ', '') if False else s for s in g['all_shapes'].tolist()]
    assert [n for n, p in model.named_parameters() if p.requires_grad] == g['trainable_names'].tolist()
    assert model.train() is None or method in ('linear', 'bitfit', 'melo')      # the overriding classes return None like the reference


@pytest.mark.skipif(not refload.available(), reason='live reference only exists in the build container')
def test_dvpt_seeded_construction_equals_reference_init():
    from oracle.cases import NEXT_CASES
    from gaviko_b200.model.dvpt import DynamicVisualPromptTuning
    ref = refload.load()
    _, kw, _ = NEXT_CASES['dvpt_t16_small']
    torch.manual_seed(3)
    a = ref.DynamicVisualPromptTuning(**kw)
    torch.manual_seed(3)
    b = DynamicVisualPromptTuning(**kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert [n for n, p in a.named_parameters() if p.requires_grad] == [n for n, p in b.named_parameters() if p.requires_grad]


@pytest.mark.skipif(not refload.available(), reason='live reference only exists in the build container')
@pytest.mark.parametrize('name', ['evp_t16_small', 'evp_mean_t16_small'])
def test_evp_seeded_construction_equals_reference_init(name):
    """EVP: nn defaults, then PromptGenerator.apply(_init_weights) with the reference's private truncated normal (model/evp.py:54-68,
    165-206) — the drop-in consumes the RNG identically, and keeps the train() quirk (frozen blocks stay in eval, the generator trains)."""
    from oracle.cases import NEXT_CASES
    from gaviko_b200.model.evp import ExplicitVisualPrompting
    ref = refload.load()
    _, kw, _ = NEXT_CASES[name]
    torch.manual_seed(3)
    a = ref.ExplicitVisualPrompting(**kw)
    torch.manual_seed(3)
    b = ExplicitVisualPrompting(**kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert [n for n, p in a.named_parameters() if p.requires_grad] == [n for n, p in b.named_parameters() if p.requires_grad]
    assert a.train() is None and b.train() is None
    for mod in ('transformer', 'conv_proj', 'dropout', 'mlp_head', 'prompt_generator'):
        assert getattr(a, mod).training == getattr(b, mod).training, mod
    a.train(False), b.train(False)
    assert a.training == b.training and a.prompt_generator.training == b.prompt_generator.training


def test_engine_refuses_trainable_backbone_tensors(tmp_path, monkeypatch):
    """--method fft, or AdaptFormer / SSF built with freeze_vit=False (their constructor default), must raise instead of silently
    returning zero gradients for backbone weights the engine has no weight-gradient kernel for; every shipped PEFT trainable set passes."""
    from oracle.cases import VARIANT_CASES
    from variant_factory import build_variant
    from gaviko_b200.model.adaptformer import AdaptFormer
    from gaviko_b200.model.vision_transformer import VisionTransformer
    monkeypatch.chdir(tmp_path)
    for name, (method, kw, _) in VARIANT_CASES.items():
        m = build_variant(method, kw)
        m._engine._check_trainable([n for n, p in m.named_parameters() if p.requires_grad])
    _, kw, _ = VARIANT_CASES['linear_t16_small']
    fft = VisionTransformer(**kw)
    with pytest.raises(NotImplementedError, match='no weight-gradient kernel'):
        fft._engine._check_trainable([n for n, p in fft.named_parameters() if p.requires_grad])
    _, kw, _ = VARIANT_CASES['adaptformer_t16_small']
    unfrozen = AdaptFormer(**dict(kw, freeze_vit=False))
    with pytest.raises(NotImplementedError):
        unfrozen._engine._check_trainable([n for n, p in unfrozen.named_parameters() if p.requires_grad])


def test_lora_cache_does_not_grow_with_optimizer_steps(tmp_path, monkeypatch):
    """the MeLO side-tensor cache used to key its entries on the parameter versions, adding four entries per layer per step."""
    from oracle.cases import VARIANT_CASES
    from variant_factory import build_variant
    monkeypatch.chdir(tmp_path)
    method, kw, _ = VARIANT_CASES['melo_t16_small']
    m = build_variant(method, kw)
    eng = m._engine
    W = eng._weights(torch.float32)
    n0 = len(eng._cache._store)
    a0 = W['layers'][0]['lora']['a_stack'].clone()
    with torch.no_grad():
        for p in m.parameters():
            if p.requires_grad:
                p.add_(0.25)                     # an optimiser step bumps every trainable tensor's version
    W = eng._weights(torch.float32)
    assert len(eng._cache._store) == n0
    assert torch.allclose(W['layers'][0]['lora']['a_stack'], a0 + 0.25)      # ... and the cached stack follows BOTH A_q and A_v
