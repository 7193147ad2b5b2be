"""Flat-buffer optimiser step for the trainable set: global-norm clip + Adam in two kernel launches, and the data-parallel
gradient exchange (one NCCL all-reduce of the flat gradient).  Semantics = reference ``src/train.py:183-189,315-319``
(``clip_grad_norm_(model.parameters(), 1.0)`` then ``Adam.step()``; ``OneCycleLR`` drives ``param_groups[0]['lr']``)."""
import ctypes as C

import torch

from . import _lib as L
from .parallel import exchange_flat_gradient


class FlatAdam(torch.optim.Optimizer):
    """Adam over one contiguous fp32 buffer.  On construction the parameters (and their .grad) are re-pointed to views of
    flat buffers, so autograd accumulates straight into the exchange buffer and no gather/scatter is needed per step."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0, process_group=None, world_size=None, model=None):
        params = list(params)
        if params and isinstance(params[0], dict):
            if len(params) != 1:
                raise ValueError('FlatAdam supports ONE parameter group (the reference trains with one, src/train.py:185-189)')
            extra = {k: v for k, v in params[0].items() if k != 'params'}
            lr, betas, eps, weight_decay = extra.get('lr', lr), extra.get('betas', betas), extra.get('eps', eps), extra.get('weight_decay', weight_decay)
            params = list(params[0]['params'])
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError('FlatAdam: no trainable parameters')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        dev = params[0].device
        if dev.type != 'cuda' or any(p.dtype != torch.float32 for p in params):
            raise L.GvkError('FlatAdam needs fp32 CUDA parameters (master weights of the trainable set)')
        self._params = params
        # every view starts on a 64-byte boundary (the kernels use float2/float4 accesses on parameters); padding stays zero
        pad = lambda k: (k + 15) // 16 * 16  # noqa: E731
        n = sum(pad(p.numel()) for p in params)
        self.numel = n
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.partials = torch.zeros(128, device=dev, dtype=torch.float32)
        self.grad_norm = torch.zeros((), device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                p.grad = self.flat_g[off:off + k].view_as(p)
                off += pad(k)
        if model is not None and hasattr(model, '_engine'):
            # let the engine's backward accumulate straight into flat_g (skips ~300 per-tensor autograd accumulations per step)
            ids = {id(p): n for n, p in model.named_parameters()}
            model._engine.grad_sink = {ids[id(p)]: p.grad for p in params if id(p) in ids}
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        # CUDA-graph mode (gaviko_b200.graph): the step counter and the learning rate live in device memory so that one captured kernel node
        # serves every replay.  None in eager mode.
        self.dyn = None
        self.group = process_group
        self.world = world_size if world_size is not None else (torch.distributed.get_world_size(process_group) if torch.distributed.is_initialized() else 1)
        if self.world > 1:
            # replicas must start identical (DDP broadcasts at construction; the reference's train.py does not seed its RNG): rank 0's values win
            torch.distributed.broadcast(self.flat_p, src=torch.distributed.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)

    def add_param_group(self, param_group):
        if getattr(self, '_params', None) is not None:
            raise ValueError('FlatAdam supports ONE parameter group')
        super().add_param_group(param_group)

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()       # grads stay views of the flat buffer
        self._realias()

    def _realias(self, gather=False):
        """`model.zero_grad()` (set_to_none=True by default) or an autograd path that replaced p.grad detaches a gradient from the flat
        buffer: fold any such stray gradient in (gather) and point p.grad at its view again."""
        off = 0
        pad = lambda k: (k + 15) // 16 * 16  # noqa: E731
        for p in self._params:
            k = p.numel()
            view = self.flat_g[off:off + k].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if gather and p.grad is not None:
                    view.add_(p.grad.to(view.dtype))
                p.grad = view
            off += pad(k)

    def state_dict(self):
        """torch.optim layout plus the flat Adam moments and the step (they live outside Optimizer.state)."""
        sd = super().state_dict()
        sd['flat_adam'] = dict(exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(), step=self.step_count, numel=self.numel)
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        fa = state_dict.pop('flat_adam', None)
        super().load_state_dict(state_dict)
        if fa is not None:
            if fa['numel'] != self.numel:
                raise ValueError('FlatAdam.load_state_dict: the trainable set changed')
            self.exp_avg.copy_(fa['exp_avg'])
            self.exp_avg_sq.copy_(fa['exp_avg_sq'])
            self.step_count = int(fa['step'])

    @torch.no_grad()
    def step(self, closure=None):
        """NB with FlatAdam the clip is part of step(): a `clip_grad_norm_` call left in the training loop would clip the LOCAL gradient before
        the all-reduce under data parallelism — remove it."""
        self._realias(gather=True)
        with torch.cuda.device(self.flat_p.device):
            self._step()
        # The update kernel writes the flat buffer behind autograd's back: bump the version counters of the parameters, so that whatever an engine
        # caches on (data_ptr, _version) of a TRAINABLE tensor (stacked LoRA factors, SSF scales folded into GEMM operands) is rebuilt next forward
        try:
            torch._C._increment_version(self._params)
        except (AttributeError, TypeError):      # private entry point moved: an in-place no-op bumps the counter as well
            for q in self._params:
                q.add_(0)

    def _step(self):
        scale = exchange_flat_gradient(self.flat_g, self.group) if self.world > 1 else 1.0   # SUM; the 1/world mean is folded into grad_scale
        g = self.param_groups[0]
        self.step_count += 1
        st = L.stream()
        L.call('gvk_grad_sumsq', C.c_void_p(self.flat_g.data_ptr()), C.c_size_t(self.numel), C.c_float(scale), C.c_void_p(self.partials.data_ptr()), st)
        if self.dyn is not None:
            step_dev, lr_dev = self.dyn
            L.call('gvk_clip_adam_dyn', C.c_void_p(self.flat_p.data_ptr()), C.c_void_p(self.flat_g.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                   C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_size_t(self.numel), C.c_void_p(self.partials.data_ptr()),
                   C.c_float(self.max_grad_norm if self.max_grad_norm else 0.0), C.c_float(scale), C.c_void_p(lr_dev.data_ptr()), C.c_float(g['betas'][0]),
                   C.c_float(g['betas'][1]), C.c_float(g['eps']), C.c_float(g['weight_decay']), C.c_void_p(step_dev.data_ptr()), C.c_void_p(self.grad_norm.data_ptr()), st)
            return
        L.call('gvk_clip_adam', C.c_void_p(self.flat_p.data_ptr()), C.c_void_p(self.flat_g.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
               C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_size_t(self.numel), C.c_void_p(self.partials.data_ptr()),
               C.c_float(self.max_grad_norm if self.max_grad_norm else 0.0), C.c_float(scale), C.c_float(g['lr']), C.c_float(g['betas'][0]),
               C.c_float(g['betas'][1]), C.c_float(g['eps']), C.c_float(g['weight_decay']), self.step_count, C.c_void_p(self.grad_norm.data_ptr()), st)
