"""Data-parallel plumbing of the hot path (one process per GPU, torch.distributed): the gradient exchange of training and the batch
sharding of inference.  Volumes are independent samples (no cross-sample operator anywhere on the path), so

* training shards the batch; the ONLY exchange per step is one SUM all-reduce of the flat trainable-gradient buffer, and the 1 / world of the
  mean is applied inside the fused clip + Adam kernel (``grad_scale``), so the global-norm clip of reference ``src/train.py:315-316`` sees the
  gradient of the global batch on every rank;
* inference shards the batch (or the file list of ``src/inference.py:141-158``) with no collective.
"""
import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def shard_range(n: int, rank: int, world: int):
    """Contiguous [begin, end) slice of n samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def exchange_flat_gradient(flat_g: torch.Tensor, group=None) -> float:
    """SUM all-reduce of the flat gradient buffer in place; returns the grad_scale (1 / world) that turns the sum of per-rank batch-mean
    gradients into the gradient of the global-batch mean (equal shard sizes, as in weak scaling)."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(flat_g, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w
