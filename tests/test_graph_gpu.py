"""GPU: the training step as one CUDA graph (gaviko_b200/graph.py) against the eager step — reference src/train.py:305-319."""
import contextlib
import io

import pytest
import torch

from gaviko_b200.graph import GraphedTrainStep
from gaviko_b200.losses.focal_loss import FocalLoss
from gaviko_b200.model.gaviko import Gaviko
from gaviko_b200.optim import FlatAdam
from oracle.cases import GAVIKO_INIT_CASES

pytestmark = pytest.mark.gpu


def _make(kw, mode, seed):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = Gaviko(**kw, compute_dtype=mode).cuda()
    return m


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_graphed_step_equals_eager_steps(mode):
    """Dropout off (eval-mode dropout modules, training graph): N replays leave the parameters where N eager steps do, with the learning rate
    changed from the host between steps (OneCycleLR's role) and fresh inputs copied into the static buffers."""
    kw, batch, seed, _ = GAVIKO_INIT_CASES['gaviko_t16_small_init']
    ma, mb = _make(kw, mode, seed), _make(kw, mode, seed)
    for m in (ma, mb):
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    oa, ob = FlatAdam(ma.parameters(), lr=1e-3, model=ma), FlatAdam(mb.parameters(), lr=1e-3, model=mb)
    crit = FocalLoss(gamma=1.2)
    g = torch.Generator().manual_seed(3)
    xs = [torch.rand(batch, kw['channels'], kw['frames'], kw['image_size'], kw['image_size'], generator=g).cuda() for _ in range(4)]
    ys = [torch.randint(0, kw['num_classes'], (batch,), generator=g).cuda() for _ in range(4)]
    lrs = [1e-3, 2e-3, 5e-4, 1e-3]
    step = GraphedTrainStep(mb, crit, ob, xs[0], ys[0])
    assert ob.step_count == 0 and torch.equal(oa.flat_p, ob.flat_p)          # the warm-up left no trace in the training state
    for x, y, lr in zip(xs, ys, lrs):
        oa.param_groups[0]['lr'] = lr
        loss_a = crit(ma(x), y)
        oa.zero_grad()
        loss_a.backward()
        oa.step()
        ob.param_groups[0]['lr'] = lr
        loss_b = step(x, y)
        assert abs(loss_a.item() - loss_b.item()) <= 1e-6 * max(1.0, abs(loss_a.item()))
    assert oa.step_count == ob.step_count == 4
    rel = ((oa.flat_p - ob.flat_p).norm() / oa.flat_p.norm()).item()
    assert rel <= 1e-6, rel


def test_graphed_step_draws_fresh_dropout_masks():
    """Dropout on: replays of one captured graph on the SAME input give different losses (the device-resident counter re-keys every mask), while
    forward and backward of a replay agree on the mask (the step still trains: the loss on a fixed batch goes down)."""
    kw, batch, seed, _ = GAVIKO_INIT_CASES['gaviko_t16_small_init']
    m = _make(dict(kw, attn_drop=0.3, proj_drop=0.3) if 'attn_drop' in kw else kw, 'bf16', seed)
    m.train()
    opt = FlatAdam(m.parameters(), lr=3e-3, model=m)
    crit = FocalLoss(gamma=1.2)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(batch, kw['channels'], kw['frames'], kw['image_size'], kw['image_size'], generator=g).cuda()
    y = torch.randint(0, kw['num_classes'], (batch,), generator=g).cuda()
    with torch.no_grad():
        m.mlp_head.head.weight.normal_(0, 0.05)       # a zero head gives the same loss whatever the masks are
    step = GraphedTrainStep(m, crit, opt, x, y)
    opt.param_groups[0]['lr'] = 0.0
    l0, l1 = step(x, y).item(), step(x, y).item()
    assert l0 != l1
    opt.param_groups[0]['lr'] = 3e-3
    losses = [step(x, y).item() for _ in range(30)]
    assert sum(losses[-5:]) / 5 < sum(losses[:5]) / 5


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_graphed_forward_equals_eager_forward(mode):
    """Inference (reference src/inference.py:105-113) as one CUDA graph: bit-identical logits to the eager forward for fresh inputs copied into the
    static buffer, and a shape the graph was not captured for is refused."""
    from gaviko_b200._lib import GvkError
    from gaviko_b200.graph import GraphedForward
    kw, batch, seed, _ = GAVIKO_INIT_CASES['gaviko_t16_small_init']
    m = _make(kw, mode, seed)
    m.eval()
    g = torch.Generator().manual_seed(5)
    xs = [torch.rand(batch, kw['channels'], kw['frames'], kw['image_size'], kw['image_size'], generator=g).cuda() for _ in range(3)]
    fwd = GraphedForward(m, xs[0])
    for x in xs[::-1]:
        with torch.no_grad():
            ref = m(x).clone()
        assert torch.equal(fwd(x), ref)
    with pytest.raises(GvkError):
        fwd(xs[0][:1])
