"""Drop-in for the reference ``src/model/adaptformer.py`` (``--method adaptformer``): a bottleneck adapter
LN -> Linear(dim, 64) -> ReLU -> Linear(64, dim) in parallel to every MLP (``adaptformer.py:22-78,93-99``).  Parameter containers
only; compute runs through ``gaviko_b200.vit_engine.VitEngine``."""
import logging
import math

import torch
from torch import nn

from ..utils.load_pretrained import load_pretrain, mapping_vit
from .vision_transformer import Attention, FeedForward, _Container, _vit_cfg, pair


class Adapter(_Container):
    def __init__(self, d_dim, down_dim=64, dropout=0.0, init_option="lora", adapter_scalar="1.0", adapter_layernorm_option="in"):
        super().__init__()
        self.d_dim = d_dim
        self.down_dim = down_dim
        self.adapter_layernorm_option = adapter_layernorm_option
        self.adapter_layer_norm_before = None
        if adapter_layernorm_option == "in" or adapter_layernorm_option == "out":
            self.adapter_layer_norm_before = nn.LayerNorm(self.d_dim)
        if adapter_scalar == "learnable_scalar":
            self.scale = nn.Parameter(torch.ones(1))
        else:
            self.scale = float(adapter_scalar)
        self.down_adapter_proj = nn.Linear(self.d_dim, self.down_dim)
        self.non_linear_func = nn.ReLU()
        self.up_adapter_proj = nn.Linear(self.down_dim, self.d_dim)
        self.dropout = dropout
        if init_option == "bert":
            raise NotImplementedError
        elif init_option == "lora":
            with torch.no_grad():
                nn.init.kaiming_uniform_(self.down_adapter_proj.weight, a=math.sqrt(5))
                nn.init.zeros_(self.up_adapter_proj.weight)
                nn.init.zeros_(self.down_adapter_proj.bias)
                nn.init.zeros_(self.up_adapter_proj.bias)


class Transformer(_Container):
    """layers.{i} = [Attention, Adapter, FeedForward] (reference adaptformer.py:81-99)."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0., adapter_dim=64):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout), Adapter(dim, down_dim=adapter_dim),
                                              FeedForward(dim, mlp_dim, dropout=dropout)]))


class AdaptFormer(nn.Module):
    def __init__(self, *, image_size, image_patch_size, frames, frame_patch_size, num_classes, pool='cls', channels=3, dim_head=64,
                 dropout=0., emb_dropout=0., backbone=None, freeze_vit=False, compute_dtype=None, adapter_dim=64, **kwargs):
        """adapter_dim: bottleneck width (the reference hard-wires 64, adaptformer.py:89; an optional kwarg here for the adapter-rank sweep of
        BASELINE.json's config 4 — the default reproduces the reference, which ignores unknown kwargs).  Multiples of 4 up to 64."""
        super().__init__()
        if adapter_dim % 4 != 0 or not 4 <= adapter_dim <= 64:
            raise ValueError('adapter_dim must be a multiple of 4 in [4, 64]')
        depth, heads, dim, mlp_dim = mapping_vit(backbone)
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(image_patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, 'Image dimensions must be divisible by the patch size.'
        assert frames % frame_patch_size == 0, 'Frames must be divisible by frame patch size'
        num_patches = (image_height // patch_height) * (image_width // patch_width) * (frames // frame_patch_size)
        self.num_patches = num_patches
        self.image_size = image_size
        self.image_patch_size = image_patch_size
        self.frames = frames
        self.frame_patch_size = frame_patch_size
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        self.conv_proj = nn.Sequential(nn.Conv3d(channels, dim, kernel_size=(frame_patch_size, image_patch_size, image_patch_size),
                                                 stride=(frame_patch_size, image_patch_size, image_patch_size)))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout, adapter_dim=adapter_dim)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Linear(dim, num_classes)
        if backbone is not None:
            logging.info(f'Loading pretrained {backbone}...')
            new_dict = load_pretrain(backbone, self.num_patches, self.conv_proj[0].weight.shape[2], './pretrained')
            self.load_state_dict(new_dict, strict=False)
            logging.info(f'Load pretrained {backbone} sucessfully!')
        self.freeze_vit = freeze_vit
        self.init_head_weights()
        if self.freeze_vit:
            for k, p in self.named_parameters():
                if "transformer" in k or "cls_token" in k or "conv_proj" in k or "pos_embedding" in k:
                    p.requires_grad = False
                if "adapter" in k or "head" in k:
                    p.requires_grad = True
        self._cfg = _vit_cfg(depth, heads, dim, mlp_dim, dim_head, channels, frames, frame_patch_size, image_height, image_width,
                             patch_height, patch_width, num_patches)
        from ..vit_engine import VitEngine
        self._engine = VitEngine(self, 'adaptformer', compute_dtype)

    def init_head_weights(self):
        nn.init.xavier_uniform_(self.mlp_head.weight)
        nn.init.zeros_(self.mlp_head.bias)
        logging.info("Initialize head weight successfully!")

    def train(self, mode=True):
        """Reference quirk preserved (adaptformer.py:176-191): returns None, train(False) only puts the children in eval mode."""
        if mode:
            super().train(mode)
            if self.freeze_vit:
                self.transformer.eval()
                self.conv_proj.eval()
                self.dropout.eval()
                self.mlp_head.train()
                for layer in self.transformer.layers:
                    layer[1].train()
        else:
            for module in self.children():
                module.eval()

    def set_compute_dtype(self, compute_dtype):
        self._engine.set_compute_dtype(compute_dtype)

    def forward(self, img):
        return self._engine(img)
