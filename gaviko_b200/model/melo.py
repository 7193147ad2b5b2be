"""Drop-in for the reference ``src/model/melo.py`` (``--method melo``): LoRA on the q and v blocks of every ``to_qkv``
(``q += (alpha // r) B_q A_q x``, ``v += (alpha // r) B_v A_v x``, ``melo.py:41-47``) around a frozen ``VisionTransformer``."""
import math

import torch
from torch import nn

from .vision_transformer import _Container


class _LoRA_qkv_timm(_Container):
    def __init__(self, qkv, linear_a_q, linear_b_q, linear_a_v, linear_b_v, r, alpha):
        super().__init__()
        self.qkv = qkv
        self.linear_a_q = linear_a_q
        self.linear_b_q = linear_b_q
        self.linear_a_v = linear_a_v
        self.linear_b_v = linear_b_v
        self.dim = qkv.in_features
        self.w_identity = torch.eye(qkv.in_features)
        self.r = r
        self.alpha = alpha


class MeLO(nn.Module):
    def __init__(self, vit, r: int, alpha: int, num_classes: int, lora_layer=None, **kwargs):
        super(MeLO, self).__init__()
        assert r > 0
        assert alpha > 0
        if lora_layer:
            self.lora_layer = lora_layer
        else:
            self.lora_layer = list(range(len(vit.transformer.layers)))
        self.w_As = []
        self.w_Bs = []
        for param in vit.parameters():
            param.requires_grad = False
        for t_layer_i, (attn, mlp) in enumerate(vit.transformer.layers):
            if t_layer_i not in self.lora_layer:
                continue
            w_qkv_linear = attn.to_qkv
            self.dim = w_qkv_linear.in_features
            w_a_linear_q = nn.Linear(self.dim, r, bias=False)
            w_b_linear_q = nn.Linear(r, self.dim, bias=False)
            w_a_linear_v = nn.Linear(self.dim, r, bias=False)
            w_b_linear_v = nn.Linear(r, self.dim, bias=False)
            self.w_As.append(w_a_linear_q)
            self.w_Bs.append(w_b_linear_q)
            self.w_As.append(w_a_linear_v)
            self.w_Bs.append(w_b_linear_v)
            attn.to_qkv = _LoRA_qkv_timm(w_qkv_linear, w_a_linear_q, w_b_linear_q, w_a_linear_v, w_b_linear_v, r, alpha)
        self.reset_parameters()
        self.lora_vit = vit
        if num_classes > 0:
            self.lora_vit.mlp_head = nn.Linear(self.dim, num_classes)
        # the wrapped VisionTransformer keeps its own engine for standalone use; this one sees the LoRA-wrapped modules and the names
        # of THIS module (``lora_vit.…``), which are the checkpoint keys (reference train.py:161-167)
        self._cfg = vit._cfg
        from ..vit_engine import VitEngine
        self._engine = VitEngine(self, 'vit', kwargs.get('compute_dtype', vit._engine._requested))

    def reset_parameters(self) -> None:
        for w_A in self.w_As:
            nn.init.kaiming_uniform_(w_A.weight, a=math.sqrt(5))
        for w_B in self.w_Bs:
            nn.init.zeros_(w_B.weight)

    def set_compute_dtype(self, compute_dtype):
        self._engine.set_compute_dtype(compute_dtype)

    def forward(self, x):
        return self._engine(x)
