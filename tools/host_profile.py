import cProfile, pstats, contextlib, io, sys, torch
sys.path.insert(0, '.')
from gaviko_b200.losses.focal_loss import FocalLoss
from gaviko_b200.model.gaviko import Gaviko
from gaviko_b200.optim import FlatAdam
from bench import GAVIKO_KW
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = Gaviko(**GAVIKO_KW, backbone='vit-b16', compute_dtype='bf16').cuda()
model.train()
opt = FlatAdam(model.parameters(), lr=1e-4, model=model)
crit = FocalLoss(gamma=1.2)
x = torch.rand(1, 1, 120, 160, 160, device='cuda'); y = torch.randint(0, 5, (1,), device='cuda')
def step():
    loss = crit(model(x), y); opt.zero_grad(); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:4500])
