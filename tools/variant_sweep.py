"""Throughput of the ViT-B/16 --method variants on the same kernels (BASELINE config 4): training step (fwd + CE + frozen-backbone bwd + clip + Adam)
in volumes/s, bf16 mode, synthetic 1x120x160x160 volumes, random-init weights.   python tools/variant_sweep.py [--batch 32] [--backbone vit-b16]"""
import argparse, contextlib, io, json, sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from variant_factory import build_variant
from gaviko_b200.losses.focal_loss import CrossEntropyLoss
from gaviko_b200.optim import FlatAdam

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=32); ap.add_argument('--backbone', default='vit-b16'); ap.add_argument('--steps', type=int, default=5); ap.add_argument('--only', default='')
a = ap.parse_args()
FULL = dict(image_size=160, image_patch_size=16, frames=120, frame_patch_size=12, num_classes=5, channels=1, pool='cls', backbone=a.backbone,
            dropout=0.1, emb_dropout=0.1, compute_dtype='bf16')       # the reference configs' dropout: active in train mode for linear / bitfit / melo
FZ = dict(FULL, freeze_vit=True)
VPT = dict(FZ, prompt_dropout=0.1, prompt_dim=64)
CASES = [('linear', FULL), ('bitfit', FULL), ('ssf', FZ), ('adaptformer', dict(FZ, adapter_dim=8)), ('adaptformer', dict(FZ, adapter_dim=16)), ('adaptformer', dict(FZ, adapter_dim=32)),
         ('adaptformer', dict(FZ, adapter_dim=64)), ('melo', dict(FULL, r=4, alpha=8)), ('melo', dict(FULL, r=8, alpha=16)),
         ('melo', dict(FULL, r=16, alpha=32)), ('shallow_vpt', dict(VPT, num_prompts=32, deep_prompt=False)), ('deep_vpt', dict(VPT, num_prompts=8, deep_prompt=True)),
         ('deep_vpt', dict(VPT, num_prompts=32, deep_prompt=True)), ('deep_vpt', dict(VPT, num_prompts=64, deep_prompt=True)),
         ('deep_vpt', dict(VPT, num_prompts=100, deep_prompt=True)), ('dvpt', dict(FZ, num_prompts=32)),
         ('evp', dict(FZ, scale_factor=4)), ('evp', dict(FZ, scale_factor=32))]      # evp.yaml ships scale_factor 4 (rank 192 at ViT-B); 32 is the constructor default
if a.only:
    CASES = [c for c in CASES if c[0] in a.only.split(',')]
x = torch.rand(a.batch, 1, 120, 160, 160, device='cuda'); y = torch.randint(0, 5, (a.batch,), device='cuda')
crit = CrossEntropyLoss()
rows = []
for method, extra in CASES:
    torch.manual_seed(0)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            m = build_variant(method, dict(extra)).cuda()
        m.train()
        opt = FlatAdam([p for p in m.parameters() if p.requires_grad], lr=1e-4, model=m)
        def step():
            loss = crit(m(x), y); opt.zero_grad(); loss.backward(); opt.step()
        for _ in range(2): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        ntr = sum(p.numel() for p in m.parameters() if p.requires_grad)
        rows.append(dict(method=method, **{k: v for k, v in extra.items() if k in ('r', 'num_prompts', 'deep_prompt', 'scale_factor', 'adapter_dim')}, ms_per_step=round(ms, 2), volumes_per_s=round(a.batch / ms * 1e3, 1), trainable=ntr))
        print(json.dumps(rows[-1]), flush=True)
    except Exception as e:   # a variant the factory cannot build with these kwargs is reported, not hidden
        print(json.dumps(dict(method=method, **{k: v for k, v in extra.items() if k in ('r', 'num_prompts', 'deep_prompt')}, error=f'{type(e).__name__}: {e}'[:200])), flush=True)
    del m
    torch.cuda.empty_cache()
