"""CPU, world_size 2 over gloo: the data-parallel exchange of the training step (gaviko_b200.parallel) reproduces the single-process gradient
of the concatenated batch, and inference sharding covers every sample exactly once.  Gradients come from the oracle (test infrastructure)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaviko_b200.parallel import exchange_flat_gradient, shard_range
from oracle import gaviko_oracle as O
from oracle.cases import GAVIKO_CASES
from oracle.golden_fill import golden_labels, golden_volume

from helpers import load_golden, sd_from_golden

NAME = 'gaviko_t16_small'


def _flat_grads(sd, names, img, y, kw):
    for n in names:
        sd[n].grad = None
    logits = O.gaviko_forward(sd, img, backbone=kw['backbone'], num_prompts=kw['num_prompts'], frame_patch_size=kw['frame_patch_size'],
                              image_patch_size=kw['image_patch_size'], local_k=kw['local_k'], DHW=kw['DHW'], share_factor=kw['share_factor'])
    O.cross_entropy(logits, y).backward()
    return torch.cat([sd[n].grad.reshape(-1) for n in names])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    kw, _ = GAVIKO_CASES[NAME]
    g = load_golden(NAME)
    sd = sd_from_golden(g)
    names = g['trainable_names'].tolist()
    batch = 4
    img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'])
    y = golden_labels(batch, kw['num_classes'])
    b0, b1 = shard_range(batch, rank, world)
    flat = _flat_grads(sd, names, img[b0:b1], y[b0:b1], kw)          # this rank's shard, loss = mean over the shard
    scale = exchange_flat_gradient(flat)
    flat *= scale
    if rank == 0:
        ref = _flat_grads(sd, names, img, y, kw)                      # single process, whole batch
        out.put(((flat - ref).norm() / ref.norm()).item())
    dist.destroy_process_group()


def test_dp2_gradient_equals_single_process_on_concatenated_batch():
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    rel = out.get()
    assert rel < 1e-5, rel


def test_shard_range_partitions_the_batch():
    for n in (1, 2, 7, 32, 33):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                b, e = shard_range(n, r, world)
                seen.extend(range(b, e))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
