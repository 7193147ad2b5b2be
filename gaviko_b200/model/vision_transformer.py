"""Parameter containers for the frozen ViT blocks (reference ``src/model/vision_transformer.py:26-72``) and the plain
``VisionTransformer`` drop-in (``:91-164``; methods linear / bitfit / melo build on it)."""
import logging

import torch
from torch import nn

from ..utils.load_pretrained import load_pretrain, mapping_vit


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError('gaviko_b200 sub-modules are parameter containers; call the top-level model')


class FeedForward(_Container):
    """net = [LayerNorm, Linear(dim, hidden), GELU, Dropout, Linear(hidden, dim), Dropout] -> keys net.0 / net.1 / net.4."""

    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Attention(_Container):
    """norm, to_qkv (no bias), to_out = [Linear, Dropout]; softmax scale dim_head**-0.5."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        project_out = not (heads == 1 and dim_head == dim)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout)) if project_out else nn.Identity()
