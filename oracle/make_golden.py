"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container (CPU, fp32).

    python -m oracle.make_golden

Weights come from ``oracle.golden_fill`` (name/shape-deterministic), inputs from ``golden_volume`` /
``golden_labels``; the files hold only outputs: logits, focal + CE loss, and every trainable gradient
under both losses.  TEST INFRASTRUCTURE ONLY.
"""
import os
import tempfile

import numpy as np
import torch

from . import refload
from .cases import EVP_INIT_CASES, GAVIKO_CASES, GAVIKO_INIT_CASES, NEXT_CASES, VARIANT_CASES
from .golden_store import chunk_sums, fingerprint, stored_in_full
from .golden_fill import golden_eval_volume, golden_fill, golden_labels, golden_volume

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def build_variant(ref, method, kw):
    if method in ('linear', 'bitfit'):
        m = ref.VisionTransformer(**kw)
        for k, v in m.named_parameters():                   # train.py:114-137
            if method == 'linear':
                v.requires_grad = 'head' in k
            else:
                v.requires_grad = ('bias' in k) or ('head' in k)
        return m
    if method == 'adaptformer':
        return ref.AdaptFormer(**kw)
    if method == 'ssf':
        return ref.ScalingShiftingFeatures(**kw)
    if method == 'melo':
        return ref.MeLO(vit=ref.VisionTransformer(**kw), **kw)   # train.py:145-147
    if method in ('deep_vpt', 'shallow_vpt'):
        return ref.PromptedVisionTransformer(**kw)
    if method == 'dvpt':
        return ref.DynamicVisualPromptTuning(**kw)
    if method == 'evp':
        return ref.ExplicitVisualPrompting(**kw)
    raise ValueError(method)


def reference_bf16_deviation(ref, model, img, y, fp32_out):
    """How far the reference's OWN pure-bf16 run (``model.to(torch.bfloat16)``, window mask cast by hand, CPU) lands from its
    fp32 run on the same weights: the bf16 noise floor our bf16 mode is judged against."""
    m = model.to(torch.bfloat16)
    for la in getattr(getattr(m, 'transformer', None), 'local_attns', []):
        la.mask = la.mask.to(torch.bfloat16)          # plain attribute, does not follow .to() (model/gaviko.py:227)
    for mod in m.modules():                            # melo.py:36 keeps an fp32 identity as a plain attribute; harmless, never used
        pass
    res = {}
    for loss_name, crit in (('focal', ref.FocalLoss(gamma=1.2)), ('ce', torch.nn.CrossEntropyLoss())):
        m.zero_grad(set_to_none=True)
        logits = m(img.bfloat16())
        crit(logits.float(), y).backward()
        num = den = 0.0
        worst = 0.0
        rows = []
        for n, p in m.named_parameters():
            if p.requires_grad:
                r = fp32_out[f'grad_{loss_name}/{n}'].astype(np.float64)
                gr = p.grad if p.grad is not None else torch.zeros_like(p)
                d = float(np.linalg.norm(gr.double().numpy() - r))
                rows.append((d, float(np.linalg.norm(r))))
                num += d * d
                den += rows[-1][1] ** 2
        gn = den ** 0.5 or 1.0          # the focal loss zeroes every gradient when all logits fall outside (1e-16, 1)
        for d, rn in rows:
            if rn > 1e-3 * gn:
                worst = max(worst, d / rn)
        res[f'refbf16_grad_global_{loss_name}'] = np.float64(num ** 0.5 / gn)
        res[f'refbf16_grad_worst_{loss_name}'] = np.float64(worst)
        lf = fp32_out['logits'].astype(np.float64)
        res['refbf16_logits_rel'] = np.float64(np.linalg.norm(logits.float().detach().numpy() - lf) / np.linalg.norm(lf))
    print('   reference bf16 deviation:', {k: float(v) for k, v in res.items()})
    return res


def run_case(ref, model, kw, batch, name, bf16_floor=False, fill=True, store='all'):
    if fill:
        golden_fill(model, seed=0)
    model.eval()
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'], channels=kw['channels'])
    y = golden_labels(batch, kw['num_classes'])
    out = {}
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    out['trainable_names'] = np.array(names)
    out['all_names'] = np.array([n for n, _ in model.named_parameters()])
    out['all_shapes'] = np.array([str(tuple(p.shape)) for _, p in model.named_parameters()])
    out['state_dict_keys'] = np.array(list(model.state_dict().keys()))
    if not fill:           # weights = the constructor's own seeded init: record a fingerprint of every tensor
        fp = fingerprint(model.state_dict())
        out['fingerprint_names'] = np.array(list(fp.keys()))
        out['fingerprint'] = np.array(list(fp.values()), dtype=np.float64)
    tr = getattr(model, 'transformer', None)
    depth = len(tr.attns) if hasattr(tr, 'attns') else (len(tr.layers) if hasattr(tr, 'layers') else 0)
    for loss_name, crit in (('focal', ref.FocalLoss(gamma=1.2)), ('ce', torch.nn.CrossEntropyLoss())):
        model.zero_grad(set_to_none=True)
        logits = model(img)
        loss = crit(logits, y)
        loss.backward()
        out['logits'] = logits.detach().numpy()
        out[f'loss_{loss_name}'] = loss.detach().numpy()
        for n, p in model.named_parameters():
            if p.requires_grad:
                assert p.grad is not None, (name, n)
                if store == 'all' or stored_in_full(n, depth):
                    out[f'grad_{loss_name}/{n}'] = p.grad.detach().numpy().copy()
                else:
                    out[f'gradsum_{loss_name}/{n}'] = chunk_sums(p.grad.detach().numpy())
                    out[f'gradnorm_{loss_name}/{n}'] = np.float64(p.grad.detach().double().norm().item())
    if bf16_floor:
        out.update(reference_bf16_deviation(ref, model, img, y, out))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'logits', out['logits'][0], 'focal', float(out['loss_focal']), 'ce', float(out['loss_ce']), 'trainable', len(names))


def focal_known_answers(ref):
    g = torch.Generator().manual_seed(5)
    cases = {}
    for i, (b, c, scale, shift) in enumerate([(7, 5, 1.0, 0.0), (4, 5, 0.3, 0.5), (6, 3, 4.0, 0.0), (5, 5, 0.2, 0.4)]):
        z = (torch.randn(b, c, generator=g) * scale + shift).requires_grad_(True)
        y = torch.randint(0, c, (b,), generator=g)
        if i == 2:
            y[1] = -100                                   # ignore_index path, focal_loss.py:101-108
        loss = ref.FocalLoss(gamma=1.2)(z, y)
        loss.backward()
        cases[f'z{i}'] = z.detach().numpy()
        cases[f'y{i}'] = y.numpy()
        cases[f'loss{i}'] = loss.detach().numpy()
        cases[f'dz{i}'] = z.grad.numpy()
    np.savez_compressed(os.path.join(OUT, 'focal_known_answers.npz'), **cases)


def window_masks(ref):
    out = {}
    for i, (dhw, k) in enumerate([((10, 10, 10), (6, 6, 6)), ((10, 10, 10), (3, 6, 6)), ((4, 4, 4), (3, 2, 2)), ((5, 4, 3), (5, 4, 2))]):
        m = ref.gaviko.LocalSelfAttention(32, local_k=k, DHW=dhw).mask[0]
        out[f'dhw{i}'] = np.array(dhw)
        out[f'k{i}'] = np.array(k)
        out[f'allow{i}'] = np.packbits((m == 0).numpy())
    np.savez_compressed(os.path.join(OUT, 'window_masks.npz'), **out)


def eval_set(ref, n=256):
    """Logits of the reference on an "eval set" of n seeded volumes (full GAViKO shape, ViT-T): the GPU test checks identical argmax."""
    kw, _ = GAVIKO_CASES['gaviko_t16_full']
    model = ref.Gaviko(**kw)
    golden_fill(model, seed=0)
    model.eval()
    out = []
    with torch.no_grad():
        for s0 in range(0, n, 8):
            img = torch.cat([golden_eval_volume(10_000 + s, kw['frames'], kw['image_size'], kw['image_size']) for s in range(s0, min(n, s0 + 8))])
            out.append(model(img))
    logits = torch.cat(out).numpy()
    top2 = np.sort(logits, 1)[:, -2:]
    print('eval set: min top-2 margin', float((top2[:, 1] - top2[:, 0]).min()), 'class histogram', np.bincount(logits.argmax(1), minlength=5))
    np.savez_compressed(os.path.join(OUT, 'gaviko_t16_full_eval256.npz'), logits=logits, seeds=np.arange(n) + 10_000)


def main():
    """python -m oracle.make_golden [case ...]   (no arguments: every golden file)"""
    import sys
    only = set(sys.argv[1:])
    want = lambda n: not only or n in only  # noqa: E731
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    ref = refload.load()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())       # vpt.py:54 appends deep_prompt.txt to the cwd
    try:
        if want('focal_known_answers'):
            focal_known_answers(ref)
        if want('window_masks'):
            window_masks(ref)
        if want('gaviko_t16_full_eval256'):
            eval_set(ref)
        for name, (kw, batch) in GAVIKO_CASES.items():
            if want(name):
                run_case(ref, ref.Gaviko(**kw), kw, batch, name, bf16_floor=True)
        for name, (kw, batch, seed, store) in GAVIKO_INIT_CASES.items():
            if want(name):
                torch.manual_seed(seed)
                model = ref.Gaviko(**kw)          # the reference's own random init (gaviko.py:445-511 + nn defaults)
                run_case(ref, model, kw, batch, name, bf16_floor=(store == 'all'), fill=False, store=store)
        for name, (method, kw, batch) in VARIANT_CASES.items():
            if want(name):
                run_case(ref, build_variant(ref, method, kw), kw, batch, name, bf16_floor=True)
        for name, (method, kw, batch) in NEXT_CASES.items():
            if want(name):
                # evp: torch.fft has no bfloat16 kernels, so the reference's own model.to(bfloat16) run does not exist (no bf16 floor recorded)
                run_case(ref, build_variant(ref, method, kw), kw, batch, name, bf16_floor=(method != 'evp'))
        for name, (kw, batch, seed, store) in EVP_INIT_CASES.items():
            if want(name):
                torch.manual_seed(seed)
                run_case(ref, ref.ExplicitVisualPrompting(**kw), kw, batch, name, bf16_floor=False, fill=False, store=store)
    finally:
        os.chdir(cwd)


if __name__ == '__main__':
    main()
