"""The training step as ONE CUDA graph: forward + loss + frozen-backbone backward + clip + Adam (reference ``src/train.py:305-319``) captured
once and replayed, so the ~560 kernel launches of a step cost one graph launch on the host instead of ~14 ms of Python + ctypes.

What changes between replays lives in device memory, because a captured kernel node keeps the arguments it was captured with:
  * the dropout seeds — every kernel with a seed mixes in a device-resident replay counter (``seed_salt`` in include/gvk.h), which the graph
    itself increments, so every replay draws fresh, forward/backward-consistent masks;
  * the Adam step counter — incremented by the graph, read by ``gvk_clip_adam_dyn``;
  * the learning rate — a device scalar filled (stream-ordered, outside the graph) from ``optimizer.param_groups[0]['lr']`` before every replay,
    so ``OneCycleLR`` (``train.py:190-198``) keeps driving it from the host.
Inputs are copied into static buffers (stream-ordered); the loss comes back as a device scalar.  Under torchrun the NCCL all-reduce of the flat
gradient is captured with the rest (every rank replays the same graph).

    step = GraphedTrainStep(model, criterion, optimizer, example_inputs, example_labels)
    for inputs, labels in loader:
        loss = step(inputs, labels)        # == criterion(model(inputs), labels); optimizer.zero_grad(); loss.backward(); optimizer.step()
        scheduler.step()
"""
import torch

from . import ops
from ._lib import GvkError
from .optim import FlatAdam


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, example_inputs, example_labels, warmup=3):
        if not isinstance(optimizer, FlatAdam):
            raise GvkError('GraphedTrainStep needs gaviko_b200.optim.FlatAdam (its clip + Adam kernels are the ones that read device-resident hyper-parameters)')
        dev = example_inputs.device
        if dev.type != 'cuda':
            raise GvkError('GraphedTrainStep needs CUDA inputs')
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.x = example_inputs.detach().clone()
        self.y = example_labels.detach().clone()
        self.counters = torch.zeros(2, device=dev, dtype=torch.int64)       # [0] dropout replay counter, [1] Adam step (1-based once incremented)
        self.counters[1] = optimizer.step_count
        self.lr = torch.zeros((), device=dev, dtype=torch.float32)
        self._one = torch.ones(2, device=dev, dtype=torch.int64)
        optimizer.dyn = (self.counters[1:2], self.lr)
        self._salt = self.counters[0:1]
        self.lr.fill_(optimizer.param_groups[0]['lr'])
        # warm-up on a side stream (lazy one-time work — function attributes, caches of frozen weights, tensor maps — must not be captured).  The
        # warm-up steps are real optimiser steps on the example batch: the training state is snapshotted and put back afterwards.
        snap = (optimizer.flat_p.clone(), optimizer.exp_avg.clone(), optimizer.exp_avg_sq.clone(), optimizer.step_count, self.counters.clone())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: under data parallelism the NCCL watchdog thread polls CUDA events while this thread captures (the all-reduce of the flat
        # gradient is captured as a graph node like any kernel)
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.loss = self._eager()
        with torch.no_grad():
            optimizer.flat_p.copy_(snap[0])
            optimizer.exp_avg.copy_(snap[1])
            optimizer.exp_avg_sq.copy_(snap[2])
            self.counters.copy_(snap[4])
        optimizer.step_count = snap[3]
        optimizer.flat_g.zero_()

    def _eager(self):
        prev = ops.SEED_SALT
        ops.SEED_SALT = self._salt
        try:
            self.counters.add_(self._one)                   # new masks, next Adam step
            loss = self.criterion(self.model(self.x), self.y)
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
        finally:
            ops.SEED_SALT = prev
        return loss.detach()

    def __call__(self, inputs, labels):
        self.x.copy_(inputs, non_blocking=True)
        self.y.copy_(labels, non_blocking=True)
        self.lr.fill_(self.optimizer.param_groups[0]['lr'])
        self.graph.replay()
        self.optimizer.step_count += 1          # mirrors the device-side counter the graph just incremented (state_dict / resume)
        return self.loss


class GraphedForward:
    """``model(inputs)`` under ``torch.no_grad()`` as ONE CUDA graph (the loop of reference ``src/inference.py:105-113`` / ``eval.py:106-114`` with a fixed
    batch shape): ~280 kernel launches per forward become one graph launch.  The model must already be in eval mode; the returned logits tensor
    is the graph's static output (copy it before the next call if it has to survive).

        fwd = GraphedForward(model, example_inputs)
        for inputs in loader:
            preds = fwd(inputs).argmax(1)
    """

    def __init__(self, model, example_inputs, warmup=2):
        dev = example_inputs.device
        if dev.type != 'cuda':
            raise GvkError('GraphedForward needs CUDA inputs')
        self.model = model
        self.x = example_inputs.detach().clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):      # lazy one-time work (weight caches, tensor maps, function attributes) must not be captured
                model(self.x)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'), torch.no_grad():
            self.out = model(self.x)

    def __call__(self, inputs):
        if inputs.shape != self.x.shape:
            raise GvkError(f'GraphedForward was captured for inputs of shape {tuple(self.x.shape)}, got {tuple(inputs.shape)}')
        self.x.copy_(inputs, non_blocking=True)
        self.graph.replay()
        return self.out
