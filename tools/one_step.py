"""Exactly N GAViKO training steps (default 2) — the command wrapped by ncu."""
import argparse, contextlib, io, sys
import torch
sys.path.insert(0, '.')
from gaviko_b200.losses.focal_loss import FocalLoss
from gaviko_b200.model.gaviko import Gaviko
from gaviko_b200.optim import FlatAdam
from bench import GAVIKO_KW
ap = argparse.ArgumentParser()
ap.add_argument('--backbone', default='vit-b16'); ap.add_argument('--batch', type=int, default=8); ap.add_argument('--dtype', default='bf16'); ap.add_argument('--steps', type=int, default=2)
a = ap.parse_args()
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = Gaviko(**GAVIKO_KW, backbone=a.backbone, compute_dtype=a.dtype).cuda()
model.train()
opt = FlatAdam(model.parameters(), lr=1e-4, model=model)
crit = FocalLoss(gamma=1.2)
x = torch.rand(a.batch, 1, 120, 160, 160, device='cuda'); y = torch.randint(0, 5, (a.batch,), device='cuda')
for _ in range(a.steps):
    loss = crit(model(x), y); opt.zero_grad(); loss.backward(); opt.step()
torch.cuda.synchronize()
print('loss', loss.item())
