"""Per-entry-point time breakdown of one GAViKO training step (CUDA events around every C-ABI call).
    python tools/profile_step.py [--backbone vit-b16] [--batch 8] [--dtype bf16]"""
import argparse, contextlib, io, sys
import torch
sys.path.insert(0, '.')
from gaviko_b200 import _lib as L
from gaviko_b200.losses.focal_loss import FocalLoss
from gaviko_b200.model.gaviko import Gaviko
from gaviko_b200.optim import FlatAdam
from bench import GAVIKO_KW

ap = argparse.ArgumentParser()
ap.add_argument('--backbone', default='vit-b16'); ap.add_argument('--batch', type=int, default=8); ap.add_argument('--dtype', default='bf16')
ap.add_argument('--steps', type=int, default=2); ap.add_argument('--by-site', action='store_true'); ap.add_argument('--eval', action='store_true')
ap.add_argument('--infer', action='store_true', help='forward only under torch.no_grad() (model.eval())')
ap.add_argument('--method', default='gaviko', help="gaviko (default) or evp (configs/evp.yaml: scale_factor 4)")
a = ap.parse_args()
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    if a.method == 'evp':
        from gaviko_b200.model.evp import ExplicitVisualPrompting
        model = ExplicitVisualPrompting(image_size=160, image_patch_size=16, frames=120, frame_patch_size=12, num_classes=5, channels=1, pool='cls', backbone=a.backbone,
                                        dropout=0.1, emb_dropout=0.1, freeze_vit=True, scale_factor=4, compute_dtype=a.dtype).cuda()
    elif a.method != 'gaviko':      # any other --method through the test factory, with the reference configs' dropout
        sys.path.insert(0, 'tests')
        from variant_factory import build_variant
        kw = dict(image_size=160, image_patch_size=16, frames=120, frame_patch_size=12, num_classes=5, channels=1, pool='cls', backbone=a.backbone, dropout=0.1,
                  emb_dropout=0.1, compute_dtype=a.dtype)
        extra = dict(melo=dict(r=4, alpha=8), adaptformer=dict(freeze_vit=True), ssf=dict(freeze_vit=True), deep_vpt=dict(freeze_vit=True, prompt_dropout=0.1, prompt_dim=64, num_prompts=32, deep_prompt=True), dvpt=dict(freeze_vit=True, num_prompts=32),
                     bitfit=dict(), linear=dict())
        model = build_variant(a.method, dict(kw, **extra.get(a.method, {}))).cuda()
    else:
        model = Gaviko(**GAVIKO_KW, backbone=a.backbone, compute_dtype=a.dtype).cuda()
model.train()
if a.eval: model.eval()
opt = FlatAdam([p for p in model.parameters() if p.requires_grad], lr=1e-4, model=model)
crit = FocalLoss(gamma=1.2)
x = torch.rand(a.batch, 1, 120, 160, 160, device='cuda'); y = torch.randint(0, 5, (a.batch,), device='cuda')

def step():
    if a.infer:
        with torch.no_grad():
            return model(x)
    loss = crit(model(x), y); opt.zero_grad(); loss.backward(); opt.step(); return loss

if a.infer:
    model.eval()

for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): step()
e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / a.steps
L.PROFILE = {}
L.PROFILE_BY_SITE = a.by_site
for _ in range(a.steps): step()
torch.cuda.synchronize()
prof, L.PROFILE = L.PROFILE, None
rows = sorted(((sum(s.elapsed_time(e) for s, e in v) / a.steps, len(v) // a.steps, k) for k, v in prof.items()), reverse=True)
acc = sum(r[0] for r in rows)
print(f'{a.backbone} {a.dtype} batch {a.batch}: {total:.2f} ms/step un-instrumented ({a.batch / total * 1e3:.1f} volumes/s); sum of bracketed calls {acc:.2f} ms; peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB')
for ms, n, k in rows:
    print(f'  {ms:9.3f} ms  {100 * ms / acc:5.1f}%  x{n:<4d} {k}')
