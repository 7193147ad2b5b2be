"""CPU: the oracle restatement reproduces the live reference's outputs (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import gaviko_oracle as O
from oracle.cases import EVP_INIT_CASES, GAVIKO_CASES, GAVIKO_INIT_CASES, NEXT_CASES, VARIANT_CASES
from oracle.golden_fill import golden_labels, golden_volume

from helpers import check_fingerprint, grad_parity, load_golden, rel_l2, sd_from_golden

TOL = 2e-5   # fp32 noise floor of the reference itself is <=6.5e-6 per tensor


def _check(g, sd, logits_fn, kw, batch):
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'], channels=kw['channels'])
    y = golden_labels(batch, kw['num_classes'])
    for loss_name, fn in (('focal', O.focal_loss), ('ce', O.cross_entropy)):
        for t in sd.values():
            t.grad = None
        logits = logits_fn(sd, img)
        loss = fn(logits, y)
        loss.backward()
        assert rel_l2(logits.detach(), g['logits']) < TOL
        assert abs(loss.item() - float(g[f'loss_{loss_name}'])) < 1e-5
        grad_parity({n: sd[n].grad for n in g['trainable_names'].tolist()}, g, loss_name, tol_global=2e-5, tol_tensor=1e-3)


@pytest.mark.parametrize('name', list(GAVIKO_CASES))
def test_gaviko_oracle_matches_reference(name):
    kw, batch = GAVIKO_CASES[name]
    g = load_golden(name)
    sd = sd_from_golden(g)
    fn = lambda sd, img: O.gaviko_forward(sd, img, backbone=kw['backbone'], num_prompts=kw['num_prompts'],
                                          frame_patch_size=kw['frame_patch_size'], image_patch_size=kw['image_patch_size'],
                                          local_k=kw['local_k'], DHW=kw['DHW'], share_factor=kw['share_factor'])
    _check(g, sd, fn, kw, batch)


@pytest.mark.parametrize('name', list(GAVIKO_INIT_CASES))
def test_gaviko_oracle_matches_reference_on_its_own_init(name):
    """The reference's own seeded random init (ViT-T, and the ViT-B / ViT-L models BASELINE.json's configs 2/3/5 name): the drop-in constructor
    reproduces the weights (fingerprint of every tensor recorded from the live reference), and the oracle restatement reproduces the
    reference's logits, losses and gradients on them."""
    import contextlib
    import io
    from gaviko_b200.model.gaviko import Gaviko
    kw, batch, seed, _ = GAVIKO_INIT_CASES[name]
    g = load_golden(name)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = Gaviko(**kw)
    check_fingerprint(m, g)
    tn = set(g['trainable_names'].tolist())
    sd = {k: v.detach().clone().requires_grad_(k in tn) for k, v in m.state_dict().items()}
    fn = lambda sd, img: O.gaviko_forward(sd, img, backbone=kw['backbone'], num_prompts=kw['num_prompts'],
                                          frame_patch_size=kw['frame_patch_size'], image_patch_size=kw['image_patch_size'],
                                          local_k=kw['local_k'], DHW=kw['DHW'], share_factor=kw['share_factor'])
    _check(g, sd, fn, kw, batch)


@pytest.mark.parametrize('name', list(VARIANT_CASES))
def test_variant_oracle_matches_reference(name):
    method, kw, batch = VARIANT_CASES[name]
    g = load_golden(name)
    sd = sd_from_golden(g)
    common = dict(backbone=kw['backbone'], frame_patch_size=kw['frame_patch_size'], image_patch_size=kw['image_patch_size'], pool=kw['pool'])
    if method in ('linear', 'bitfit'):
        fn = lambda sd, img: O.vit_forward(sd, img, **common)
    elif method == 'adaptformer':
        fn = lambda sd, img: O.adaptformer_forward(sd, img, **common)
    elif method == 'melo':
        fn = lambda sd, img: O.melo_forward(sd, img, r=kw['r'], alpha=kw['alpha'], **common)
    elif method == 'ssf':
        fn = lambda sd, img: O.ssf_forward(sd, img, **common)
    else:
        fn = lambda sd, img: O.vpt_forward(sd, img, deep_prompt=kw['deep_prompt'], **common)
    _check(g, sd, fn, kw, batch)


@pytest.mark.parametrize('name', list(NEXT_CASES))
def test_next_method_oracle_matches_reference(name):
    """SURVEY.md §8 (f3) DVPT and (f4) EVP: the oracle restatements against the live reference (DVPT in both pool modes, trainable set =
    prompts + every prompt_proj tensor + head; EVP at scale_factor 4 / 32, trainable set = every prompt_generator tensor + head)."""
    method, kw, batch = NEXT_CASES[name]
    g = load_golden(name)
    sd = sd_from_golden(g)
    names = g['trainable_names'].tolist()
    common = dict(backbone=kw['backbone'], frame_patch_size=kw['frame_patch_size'], image_patch_size=kw['image_patch_size'], pool=kw['pool'])
    if method == 'dvpt':
        assert 'prompt_embeddings' in names and 'prompt_positional_embedding' in names and 'mlp_head.weight' in names
        assert all(('prompt' in n) or ('head' in n) for n in names) and len(names) == 2 + 5 * 12 + 2
        fn = lambda sd, img: O.dvpt_forward(sd, img, num_prompts=kw['num_prompts'], **common)
    else:
        assert all(n.startswith('prompt_generator.') or n.startswith('mlp_head.') for n in names) and len(names) == 2 + 4 + 2 * 12 + 2
        fn = lambda sd, img: O.evp_forward(sd, img, freq_nums=kw['freq_nums'], **common)
    _check(g, sd, fn, kw, batch)


@pytest.mark.parametrize('name', list(EVP_INIT_CASES))
def test_evp_oracle_matches_reference_on_its_own_init(name):
    """EVP at the shipped geometry / backbone (ViT-B, scale_factor 4, full volumes) on the reference's own seeded init: the drop-in constructor
    reproduces the weights (fingerprint recorded from the live reference) and the oracle reproduces logits, losses and gradients on them."""
    from gaviko_b200.model.evp import ExplicitVisualPrompting
    kw, batch, seed, _ = EVP_INIT_CASES[name]
    g = load_golden(name)
    torch.manual_seed(seed)
    m = ExplicitVisualPrompting(**kw)
    check_fingerprint(m, g)
    tn = set(g['trainable_names'].tolist())
    sd = {k: v.detach().clone().requires_grad_(k in tn) for k, v in m.state_dict().items()}
    fn = lambda sd, img: O.evp_forward(sd, img, backbone=kw['backbone'], frame_patch_size=kw['frame_patch_size'], image_patch_size=kw['image_patch_size'],
                                       pool=kw['pool'], freq_nums=kw['freq_nums'])
    _check(g, sd, fn, kw, batch)


@pytest.mark.parametrize('shape,rate', [((2, 1, 48, 64, 64), 0.25), ((1, 1, 24, 32, 48), 0.1), ((2, 1, 12, 32, 32), 0.9), ((1, 2, 20, 16, 24), 0.5),
                                        ((1, 1, 120, 160, 160), 0.25)])
def test_evp_highpass_closed_form(shape, rate):
    """PromptGenerator.fft (model/evp.py:124-146, restated with the same torch.fft calls) equals its closed form — a real H x H matrix applied
    along H on a cyclic range of depth slices, |x| elsewhere — which is what gvk_hfreq_filter computes."""
    img = torch.rand(shape, generator=torch.Generator().manual_seed(3))
    a, b = O.evp_highpass(img, rate), O.evp_highpass_closed_form(img, rate)
    assert (a - b).abs().max().item() < 2e-6
    d_un, k_un = O.evp_filter_plan(shape, rate)
    if shape[2:] == (120, 160, 160):      # the shipped geometry (configs/evp.yaml): 80 of 120 slices lose the 80 lowest H-frequencies
        assert int(d_un.sum()) == 80 and int(k_un.sum()) == 80 and bool(k_un[0]) and not bool(d_un[60])


def test_focal_known_answers():
    g = load_golden('focal_known_answers')
    for i in range(4):
        z = torch.tensor(g[f'z{i}'], requires_grad=True)
        y = torch.tensor(g[f'y{i}'])
        loss = O.focal_loss(z, y)
        loss.backward()
        assert loss.item() == pytest.approx(float(g[f'loss{i}']), rel=1e-6)
        np.testing.assert_allclose(z.grad.numpy(), g[f'dz{i}'], rtol=1e-5, atol=1e-8)


def test_window_mask_formula():
    g = load_golden('window_masks')
    for i in range(4):
        allow = O.window_allow(tuple(g[f'dhw{i}']), tuple(g[f'k{i}']))
        assert np.array_equal(np.packbits(allow.numpy()), g[f'allow{i}'])
    a = O.window_allow((10, 10, 10), (6, 6, 6))
    assert int(a.sum()) == 132651 and int(a.sum(1).min()) == 27 and int(a.sum(1).max()) == 216


def test_eval_set_oracle_matches_reference():
    """The eval-set golden (reference logits of 256 structured volumes, eval mode) against the oracle on a few of its volumes."""
    from oracle.golden_fill import golden_eval_volume
    g = load_golden('gaviko_t16_full_eval256')
    kw, _ = GAVIKO_CASES['gaviko_t16_full']
    sd = sd_from_golden(load_golden('gaviko_t16_full'))
    pick = [0, 101, 255]
    img = torch.cat([golden_eval_volume(int(g['seeds'][i]), kw['frames'], kw['image_size'], kw['image_size']) for i in pick])
    with torch.no_grad():
        logits = O.gaviko_forward(sd, img, backbone=kw['backbone'], num_prompts=kw['num_prompts'], frame_patch_size=kw['frame_patch_size'],
                                  image_patch_size=kw['image_patch_size'], local_k=kw['local_k'], DHW=kw['DHW'], share_factor=kw['share_factor'])
    assert rel_l2(logits, g['logits'][pick]) < TOL
    assert (logits.argmax(1).numpy() == g['logits'][pick].argmax(1)).all()
