"""GPU parity of the drop-in Gaviko module against golden outputs of the live reference (tests/golden, oracle/make_golden.py)
and against the oracle restatement on the same weights, through the public module API (-> C ABI -> sm_100a kernels)."""
import pytest
import torch

from gaviko_b200.losses.focal_loss import CrossEntropyLoss, FocalLoss
from gaviko_b200.model.gaviko import Gaviko
from oracle.cases import GAVIKO_CASES, GAVIKO_INIT_CASES
from oracle.golden_fill import golden_fill, golden_labels, golden_volume

from helpers import check_fingerprint, grad_parity, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _build(name, compute_dtype):
    kw, batch = GAVIKO_CASES[name]
    model = Gaviko(**kw, compute_dtype=compute_dtype)
    golden_fill(model, seed=0)
    model = model.cuda()
    model.eval()                      # dropout off for parity (reference quirk: self.training stays True)
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'], channels=kw['channels']).cuda()
    y = golden_labels(batch, kw['num_classes']).cuda()
    return model, img, y


@pytest.mark.parametrize('name', list(GAVIKO_CASES))
def test_gaviko_fp32_matches_reference(name):
    """fp32 mode: logits and every trainable gradient within 1e-4 relative of the reference (north star tolerance)."""
    g = load_golden(name)
    model, img, y = _build(name, 'fp32')
    assert [n for n, p in model.named_parameters() if p.requires_grad] == g['trainable_names'].tolist()
    for loss_name, crit in (('focal', FocalLoss(gamma=1.2)), ('ce', CrossEntropyLoss())):
        model.zero_grad(set_to_none=True)
        logits = model(img)
        loss = crit(logits, y)
        loss.backward()
        assert rel_l2(logits.detach().cpu(), g['logits']) < 1e-4, rel_l2(logits.detach().cpu(), g['logits'])
        assert abs(loss.item() - float(g[f'loss_{loss_name}'])) < 1e-4
        grads = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
        glob, worst, wname = grad_parity(grads, g, loss_name, tol_global=1e-4, tol_tensor=1e-3)
        print(f'{name} {loss_name}: logits rel {rel_l2(logits.detach().cpu(), g["logits"]):.2e} grads global {glob:.2e} worst {worst:.2e} ({wname})')


# bf16-mode gradient error of the golden_fill cases, measured on a B200 (profiles/parity_r02.txt) and reproduced by the CPU emulation of every
# rounding site (tools/error_budget.py -> profiles/error_budget_r02.txt).  golden_fill draws every Linear weight with std fan_in^-0.5, 1.73x the
# reference's own init, which makes the 12-layer network ill-conditioned: rounding ONE frozen weight matrix family to bf16 already moves the
# gradient by 2-5e-2 on the small case, and all backward roundings together contribute < 1e-3.  These cases therefore assert the measured floor
# x 1.2; the north star's 2e-2 is asserted on the reference's own random init below (test_gaviko_init_cases_bf16).
_BF16_FLOOR = {('gaviko_t16_full', 'focal'): 2.25e-2, ('gaviko_t16_full', 'ce'): 3.6e-2, ('gaviko_t16_small', 'focal'): 8.7e-2, ('gaviko_t16_small', 'ce'): 8.8e-2}
_BF16_FLOOR_TENSOR = {'gaviko_t16_full': 9.4e-2, 'gaviko_t16_small': 1.66e-1}


@pytest.mark.parametrize('name', list(GAVIKO_CASES))
def test_gaviko_bf16_matches_reference(name):
    """bf16 mode on the golden_fill weights: logits within 2e-2 relative with identical argmax; gradients within 1.2x the measured floor of
    bf16-operand arithmetic on these weights (see _BF16_FLOOR), and never further from the fp32 reference than the reference's own bf16 run."""
    g = load_golden(name)
    model, img, y = _build(name, 'bf16')
    for loss_name, crit in (('focal', FocalLoss(gamma=1.2)), ('ce', CrossEntropyLoss())):
        model.zero_grad(set_to_none=True)
        logits = model(img)
        loss = crit(logits, y)
        loss.backward()
        rl = rel_l2(logits.detach().cpu(), g['logits'])
        assert rl < 2e-2, rl
        assert logits.argmax(1).cpu().tolist() == g['logits'].argmax(1).tolist()
        grads = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
        tol_g = 1.2 * _BF16_FLOOR[(name, loss_name)]
        assert tol_g < float(g[f'refbf16_grad_global_{loss_name}'])
        # sub-floor tensors (< 1e-3 of the global norm, e.g. gl_balancer gates = sums of ctx_g - ctx_l differences) are cancellation noise in
        # any 8-bit-mantissa run (measured worst: layer-11 gl_balancer weight, 2.4e-4 of the global norm, off by 2.7e-4 of it under CE); they may
        # each add at most 3 * tol_t * 1e-3 = 3.4e-4 of the global norm
        glob, worst, wname = grad_parity(grads, g, loss_name, tol_global=tol_g, tol_tensor=1.2 * _BF16_FLOOR_TENSOR[name], floor=1e-3, floor_slack=3.0)
        print(f'{name} {loss_name}: logits rel {rl:.2e} grads global {glob:.2e} worst {worst:.2e} ({wname})')


def _build_init(name, compute_dtype):
    import contextlib
    import io
    kw, batch, seed, _ = GAVIKO_INIT_CASES[name]
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = Gaviko(**kw, compute_dtype=compute_dtype)
    g = load_golden(name)
    check_fingerprint(model, g)          # same weights as the reference built under this seed
    model = model.cuda()
    model.eval()
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'], channels=kw['channels']).cuda()
    y = golden_labels(batch, kw['num_classes']).cuda()
    return model, img, y, g


def _run_losses(model, img, y):
    for loss_name, crit in (('focal', FocalLoss(gamma=1.2)), ('ce', CrossEntropyLoss())):
        model.zero_grad(set_to_none=True)
        logits = model(img)
        loss = crit(logits, y)
        loss.backward()
        yield loss_name, logits, loss, {n: p.grad for n, p in model.named_parameters() if p.requires_grad}


@pytest.mark.parametrize('name', list(GAVIKO_INIT_CASES))
def test_gaviko_init_cases_fp32(name):
    """The north star's setting — the reference's own random-init weights, ViT-T / ViT-B (the model BASELINE.json's metric is quoted on) / ViT-L:
    fp32 mode within 1e-4 relative on logits, loss and (globally) on the gradients."""
    model, img, y, g = _build_init(name, 'fp32')
    for loss_name, logits, loss, grads in _run_losses(model, img, y):
        rl = rel_l2(logits.detach().cpu(), g['logits'])
        assert rl < 1e-4, rl
        assert abs(loss.item() - float(g[f'loss_{loss_name}'])) < 1e-4
        glob, worst, wname = grad_parity(grads, g, loss_name, tol_global=1e-4, tol_tensor=1e-3)
        print(f'{name} fp32 {loss_name}: logits rel {rl:.2e} grads global {glob:.2e} worst {worst:.2e} ({wname})')


@pytest.mark.parametrize('name', list(GAVIKO_INIT_CASES))
def test_gaviko_init_cases_bf16(name):
    """bf16 mode on the reference's own random-init weights: logits and gradients within the north star's 2e-2 relative (global rel-L2), identical
    argmax.  Per tensor (above 1e-3 of the global norm): 6e-2, the spread bf16 operand rounding leaves on single small tensors (emulated floor
    2.2e-2 .. 5e-2, tools/error_budget.py).  The focal loss has exactly zero gradient on these logits (all outside (1e-16, 1)): the
    CUDA path must return exact zeros there."""
    model, img, y, g = _build_init(name, 'bf16')
    for loss_name, logits, loss, grads in _run_losses(model, img, y):
        rl = rel_l2(logits.detach().cpu(), g['logits'])
        assert rl < 2e-2, rl
        assert logits.argmax(1).cpu().tolist() == g['logits'].argmax(1).tolist()
        glob, worst, wname = grad_parity(grads, g, loss_name, tol_global=2e-2, tol_tensor=6e-2, floor=1e-3, floor_slack=2.0)
        print(f'{name} bf16 {loss_name}: logits rel {rl:.2e} grads global {glob:.2e} worst {worst:.2e} ({wname})')


def test_inference_no_grad_and_determinism():
    model, img, y = _build('gaviko_t16_small', 'fp32')
    with torch.no_grad():
        a = model(img)
        b = model(img)
    assert torch.equal(a, b) and not a.requires_grad


def test_train_mode_dropout_is_active_and_unbiased():
    kw, batch = GAVIKO_CASES['gaviko_t16_small']
    kw = dict(kw, attn_drop=0.2, proj_drop=0.2)
    model = Gaviko(**kw, compute_dtype='fp32')
    golden_fill(model, seed=0)
    model = model.cuda()
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size']).cuda()
    model.eval()
    with torch.no_grad():
        ref = model(img)
    assert model.train() is None       # reference quirk (model/gaviko.py:513-528)
    with torch.no_grad():
        outs = torch.stack([model(img) for _ in range(8)])
    assert (outs[0] != outs[1]).any()                      # masks differ between steps
    assert (outs.mean(0) - ref).abs().max().item() < 0.5   # and do not change the scale
    logits = model(img)
    FocalLoss(gamma=1.2)(logits, golden_labels(batch, 5).cuda()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters() if p.requires_grad)


def test_rejects_cpu_input():
    from gaviko_b200._lib import GvkError
    kw, batch = GAVIKO_CASES['gaviko_t16_small']
    model = Gaviko(**kw)
    with pytest.raises(GvkError):
        model(torch.zeros(1, 1, 48, 64, 64))


def test_flat_adam_sink_and_step_match_torch():
    """FlatAdam (engine accumulates straight into the flat buffer; fused clip+Adam kernels) == autograd grads + torch clip + torch Adam.
    Gradients are compared every step; the torch side then steps on a copy of OUR gradients, so the parameter comparison isolates the
    optimiser kernels (Adam's 1/sqrt(v) normalisation would otherwise amplify the atomics-order noise of near-zero gradient elements)."""
    from gaviko_b200.optim import FlatAdam
    ma, img, y = _build('gaviko_t16_small', 'fp32')
    mb, _, _ = _build('gaviko_t16_small', 'fp32')
    crit = CrossEntropyLoss()
    opt_a = FlatAdam(ma.parameters(), lr=1e-3, eps=1e-8, max_grad_norm=1.0, model=ma)
    tb = [p for p in mb.parameters() if p.requires_grad]
    opt_b = torch.optim.Adam(tb, lr=1e-3, eps=1e-8)
    for step in range(3):
        opt_a.zero_grad()
        crit(ma(img), y).backward()
        opt_b.zero_grad()
        crit(mb(img), y).backward()
        for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            if pa.requires_grad:
                scale = pb.grad.abs().max().item()
                assert (pa.grad - pb.grad).abs().max().item() <= 1e-3 * scale + 1e-9, (step, n)
                pb.grad.copy_(pa.grad)
        norm_b = torch.nn.utils.clip_grad_norm_(mb.parameters(), 1.0)
        opt_a.step()
        opt_b.step()
        assert abs(opt_a.grad_norm.item() - norm_b.item()) <= 1e-5 * norm_b.item()
        for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            if pa.requires_grad:
                assert torch.allclose(pa, pb, rtol=1e-5, atol=2e-7), (step, n, (pa - pb).abs().max().item())


def test_flat_adam_survives_model_zero_grad_and_round_trips_its_state():
    """`model.zero_grad()` (set_to_none=True) detaches p.grad from the flat buffer — step() must re-alias it (and fold in a gradient
    autograd allocated meanwhile); state_dict() / load_state_dict() carry the Adam moments and the step; one parameter group only."""
    from gaviko_b200.optim import FlatAdam
    ma, img, y = _build('gaviko_t16_small', 'fp32')
    mb, _, _ = _build('gaviko_t16_small', 'fp32')
    crit = CrossEntropyLoss()
    opt_a = FlatAdam(ma.parameters(), lr=1e-3, model=ma)
    opt_b = FlatAdam(mb.parameters(), lr=1e-3, model=mb)
    for step in range(2):
        ma.zero_grad()                       # torch default: set_to_none=True
        assert all(p.grad is None for p in ma.parameters() if p.requires_grad)
        opt_a.zero_grad()
        crit(ma(img), y).backward()
        opt_a.step()
        opt_b.zero_grad()
        crit(mb(img), y).backward()
        opt_b.step()
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):      # equal up to the atomics-order noise of the fusion gradients
        assert torch.allclose(pa, pb, rtol=1e-4, atol=1e-6), (n, (pa - pb).abs().max().item())
    # a stray gradient (not a view of the flat buffer) is folded in
    mb._engine.grad_sink = None
    opt_b.flat_g.zero_()
    for p in mb.parameters():
        p.grad = None
    crit(mb(img), y).backward()
    stray = {n: p.grad.clone() for n, p in mb.named_parameters() if p.requires_grad}
    opt_b._realias(gather=True)
    for n, p in mb.named_parameters():
        if p.requires_grad:
            assert p.grad.data_ptr() >= opt_b.flat_g.data_ptr() and torch.equal(p.grad, stray[n]), n
    sd = opt_a.state_dict()
    assert sd['flat_adam']['step'] == 2 and sd['flat_adam']['exp_avg'].abs().sum().item() > 0
    mc, _, _ = _build('gaviko_t16_small', 'fp32')
    opt_c = FlatAdam(mc.parameters(), lr=1e-3, model=mc)
    opt_c.load_state_dict(sd)
    assert opt_c.step_count == 2 and torch.equal(opt_c.exp_avg, opt_a.exp_avg) and torch.equal(opt_c.exp_avg_sq, opt_a.exp_avg_sq)
    with pytest.raises(ValueError):
        FlatAdam([dict(params=[p for p in mc.parameters() if p.requires_grad][:2]), dict(params=[p for p in mc.parameters() if p.requires_grad][2:])])


def test_loss_poisons_out_of_range_labels():
    """an out-of-range label used to index past the probability row; the reference raises, the kernel returns NaN."""
    z = torch.randn(4, 5, device='cuda', requires_grad=True)
    y = torch.tensor([0, 4, 7, 1], device='cuda')
    for crit in (FocalLoss(gamma=1.2), CrossEntropyLoss()):
        loss = crit(z, y)
        assert torch.isnan(loss).item()
    y_ok = torch.tensor([0, 4, -100, 1], device='cuda')
    assert torch.isfinite(FocalLoss(gamma=1.2)(z, y_ok)).item()


@pytest.mark.parametrize('mode,margin', [('fp32', 1e-3), ('bf16', 5e-2)])
def test_eval_set_argmax_identical(mode, margin):
    """North star: identical argmax predictions on the eval set.  256 structured synthetic volumes (oracle.golden_fill.golden_eval_volume) run
    through the reference (golden logits from oracle/make_golden.py) and through the CUDA path.  With random-init weights the raw argmax is the
    same class for every volume, so the test also compares the argmax of the logits centred on the reference's per-class means — the part that
    depends on the input — for every volume whose reference top-2 margin there exceeds `margin`."""
    import numpy as np
    from oracle.golden_fill import golden_eval_volume
    g = load_golden('gaviko_t16_full_eval256')
    ref = g['logits']
    kw, _ = GAVIKO_CASES['gaviko_t16_full']
    model = Gaviko(**kw, compute_dtype=mode)
    golden_fill(model, seed=0)
    model = model.cuda()
    model.eval()
    outs = []
    with torch.no_grad():
        for s0 in range(0, len(ref), 32):
            img = torch.cat([golden_eval_volume(int(s), kw['frames'], kw['image_size'], kw['image_size']) for s in g['seeds'][s0:s0 + 32]]).cuda()
            outs.append(model(img).float().cpu())
    got = torch.cat(outs).numpy()
    assert (got.argmax(1) == ref.argmax(1)).all()
    mean = ref.mean(0)
    cr, cg = ref - mean, got - mean
    top2 = np.sort(cr, 1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > margin
    assert decided.mean() > 0.5
    assert (cg.argmax(1)[decided] == cr.argmax(1)[decided]).all(), int((cg.argmax(1)[decided] != cr.argmax(1)[decided]).sum())
    print(f'{mode}: max |logit diff| {np.abs(got - ref).max():.2e}; centred argmax compared on {int(decided.sum())} of {len(ref)} volumes')


def test_trainable_only_checkpoint_roundtrip(tmp_path):
    """Checkpoint contract of reference train.py:161-167,478-483 and eval.py:87-92: save {k: v for k in tuning_params}, load it with
    strict=False over a freshly built model holding the same frozen backbone, get identical logits."""
    ma, img, y = _build('gaviko_t16_small', 'fp32')
    tuning = [n for n, p in ma.named_parameters() if p.requires_grad]
    opt = torch.optim.Adam([p for p in ma.parameters() if p.requires_grad], lr=1e-2)
    for _ in range(2):                                   # move the trainables away from their initial values
        opt.zero_grad()
        CrossEntropyLoss()(ma(img), y).backward()
        torch.nn.utils.clip_grad_norm_(ma.parameters(), 1.0)
        opt.step()
    ckpt = {k: v for k, v in ma.state_dict().items() if k in tuning}
    assert sorted(ckpt) == sorted(tuning)
    path = tmp_path / 'gaviko_vit-t16_best_model.pt'
    torch.save(ckpt, path)
    mb, _, _ = _build('gaviko_t16_small', 'fp32')        # same frozen backbone (golden_fill), initial trainables
    with torch.no_grad():
        before = mb(img)
        want = ma(img)
    assert not torch.allclose(before, want, atol=1e-4)
    missing, unexpected = mb.load_state_dict(torch.load(path), strict=False)
    assert not unexpected and all(k not in tuning for k in missing)
    with torch.no_grad():
        assert torch.equal(mb(img), want)
