"""Print the parity of the drop-in Gaviko module against the golden reference outputs (no assertions): per mode and loss,
logits rel-L2, global gradient rel-L2, the distribution of per-tensor rel-L2 and the worst tensors.
    python tools/parity_report.py [case ...] > profiles/parity_rNN.txt
"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from gaviko_b200.losses.focal_loss import CrossEntropyLoss, FocalLoss  # noqa: E402
from gaviko_b200.model.gaviko import Gaviko  # noqa: E402
from helpers import load_golden, rel_l2  # noqa: E402
from oracle.cases import GAVIKO_CASES, GAVIKO_INIT_CASES  # noqa: E402
from oracle.golden_fill import golden_fill, golden_labels, golden_volume  # noqa: E402

import contextlib, io  # noqa: E402

from oracle.golden_store import chunk_sums  # noqa: E402

cases = sys.argv[1:] or (list(GAVIKO_CASES) + list(GAVIKO_INIT_CASES))
for name in cases:
    init = name in GAVIKO_INIT_CASES
    kw, batch = GAVIKO_INIT_CASES[name][:2] if init else GAVIKO_CASES[name]
    g = load_golden(name)
    for mode in ('fp32', 'bf16'):
        if init:
            torch.manual_seed(GAVIKO_INIT_CASES[name][2])
        with contextlib.redirect_stdout(io.StringIO()):
            model = Gaviko(**kw, compute_dtype=mode)
        if not init:
            golden_fill(model, seed=0)
        model = model.cuda()
        model.eval()
        img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size']).cuda()
        y = golden_labels(batch, kw['num_classes']).cuda()
        for loss_name, crit in (('focal', FocalLoss(gamma=1.2)), ('ce', CrossEntropyLoss())):
            model.zero_grad(set_to_none=True)
            logits = model(img)
            loss = crit(logits, y)
            loss.backward()
            rows, num, den = [], 0.0, 0.0
            for n, p in model.named_parameters():
                if not p.requires_grad:
                    continue
                a = p.grad.double().cpu().numpy()
                if f'grad_{loss_name}/{n}' in g.files:
                    r = g[f'grad_{loss_name}/{n}'].astype(np.float64)
                else:                                   # 'subset' store: 32-element chunk sums
                    r, a = g[f'gradsum_{loss_name}/{n}'].astype(np.float64), chunk_sums(a)
                d = float(np.linalg.norm(a - r))
                rn = float(np.linalg.norm(r))
                num += d * d
                den += rn * rn
                rows.append((d / rn if rn > 0 else (0.0 if d == 0 else float('inf')), rn, n))
            gn = den ** 0.5
            rels = np.array([r[0] for r in rows if np.isfinite(r[0])])
            sig = np.array([r[0] for r in rows if r[1] > 1e-3 * gn])
            print(f'[{name} | {mode} | {loss_name}] logits rel-L2 {rel_l2(logits.detach().cpu(), g["logits"]):.3e}  argmax_equal {logits.argmax(1).cpu().tolist() == g["logits"].argmax(1).tolist()}  '
                  f'loss {loss.item():.6f} (ref {float(g["loss_" + loss_name]):.6f})')
            print(f'    grads: global rel-L2 {(num ** 0.5) / (gn or 1):.3e} | per-tensor median {np.median(rels):.3e} p90 {np.percentile(rels, 90):.3e} max {rels.max():.3e} | '
                  f'tensors with norm > 1e-3*global: {len(sig)} of {len(rows)}, max rel among them {sig.max() if len(sig) else 0:.3e}')
            for rel, rn, n in sorted(rows, reverse=True)[:5]:
                print(f'      {rel:.3e}  |ref|={rn:.3e} ({rn / (gn or 1):.1e} of global)  {n}')
        del model
