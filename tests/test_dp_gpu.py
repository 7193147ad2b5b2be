"""GPU, two ranks over NCCL (skipped on a single-GPU box): one data-parallel training step of the CUDA path — per-rank shards, the flat-gradient
all-reduce inside FlatAdam, clip on the reduced gradient, Adam — leaves both ranks with bit-identical parameters that equal a single-process
step on the concatenated batch (reference step semantics train.py:305-319)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

NAME = 'gaviko_t16_small'


def _one_step(model, img, y, world_size):
    from gaviko_b200.losses.focal_loss import CrossEntropyLoss
    from gaviko_b200.optim import FlatAdam
    opt = FlatAdam(model.parameters(), lr=1e-2, max_grad_norm=1.0, model=model, world_size=world_size)
    opt.zero_grad()
    CrossEntropyLoss()(model(img), y).backward()
    opt.step()
    return opt.flat_p.clone(), opt.grad_norm.item()


def _worker(rank, world, port, out):
    from gaviko_b200.model.gaviko import Gaviko
    from gaviko_b200.parallel import shard_range
    from oracle.cases import GAVIKO_CASES
    from oracle.golden_fill import golden_fill, golden_labels, golden_volume
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    kw, _ = GAVIKO_CASES[NAME]

    def build():
        m = Gaviko(**kw, compute_dtype='fp32')
        golden_fill(m, seed=0)
        m = m.cuda()
        m.eval()                              # dropout off: the two runs must see the same arithmetic
        return m

    batch = 4
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size']).cuda()
    y = golden_labels(batch, kw['num_classes']).cuda()
    b0, b1 = shard_range(batch, rank, world)
    mine, gnorm = _one_step(build(), img[b0:b1], y[b0:b1], world)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        ref, gnorm_ref = _one_step(build(), img, y, 1)
        out.put((all(torch.equal(gathered[0], g) for g in gathered), ((mine - ref).norm() / ref.norm()).item(), abs(gnorm - gnorm_ref) / gnorm_ref,
                 ((mine - ref).abs().max() / 1e-2).item()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_dp2_step_matches_single_gpu_step_on_the_concatenated_batch():
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    identical, rel, gnorm_rel, max_step_frac = out.get()
    assert identical                        # every rank holds the same parameters after the step
    assert rel < 1e-6 and gnorm_rel < 1e-4  # and they are the single-GPU parameters (Adam's first step moves each weight by ~lr: compare in units of lr)
    assert max_step_frac < 1e-2, max_step_frac
    print(f'dp2 vs single GPU: params rel-L2 {rel:.2e}, clip norm rel {gnorm_rel:.2e}, max |delta| {max_step_frac:.2e} of lr')
