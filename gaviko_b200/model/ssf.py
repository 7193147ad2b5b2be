"""Drop-in for the reference ``src/model/ssf.py`` (``--method ssf``): a trainable ``x * scale + shift`` after the patch embedding,
every LayerNorm and every Linear (``ssf.py:24-31,64-116,133-138,236``)."""
import logging

import torch
from torch import nn

from ..utils.load_pretrained import load_pretrain, mapping_vit
from .vision_transformer import _Container, _vit_cfg, pair


def init_ssf_scale_shift(dim):
    scale = nn.Parameter(torch.ones(dim))
    shift = nn.Parameter(torch.zeros(dim))
    nn.init.normal_(scale, mean=1, std=.02)
    nn.init.normal_(shift, std=.02)
    return scale, shift


class FeedForward(_Container):
    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.ssf_scale_0, self.ssf_shift_0 = init_ssf_scale_shift(dim)
        self.ssf_scale_1, self.ssf_shift_1 = init_ssf_scale_shift(hidden_dim)
        self.ssf_scale_2, self.ssf_shift_2 = init_ssf_scale_shift(dim)
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Attention(_Container):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        project_out = not (heads == 1 and dim_head == dim)
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout)) if project_out else nn.Identity()
        self.ssf_scale_0, self.ssf_shift_0 = init_ssf_scale_shift(dim)
        self.ssf_scale_1, self.ssf_shift_1 = init_ssf_scale_shift(inner_dim * 3)
        self.ssf_scale_2, self.ssf_shift_2 = init_ssf_scale_shift(dim)


class Transformer(_Container):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.):
        super().__init__()
        self.ls1 = nn.Identity()
        self.ls2 = nn.Identity()
        self.norm = nn.LayerNorm(dim)
        self.ssf_scale_1, self.ssf_shift_1 = init_ssf_scale_shift(dim)
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout),
                                              FeedForward(dim, mlp_dim, dropout=dropout)]))


class ScalingShiftingFeatures(nn.Module):
    def __init__(self, *, image_size, image_patch_size, frames, frame_patch_size, num_classes, pool='cls', channels=3, dim_head=64,
                 dropout=0., emb_dropout=0., backbone=None, freeze_vit=False, compute_dtype=None, **kwargs):
        super().__init__()
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
        self.ssf_scale_1, self.ssf_shift_1 = init_ssf_scale_shift(dim)
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Linear(dim, num_classes)
        self.freeze_vit = freeze_vit
        self.init_head_weights()
        if backbone is not None:
            logging.info(f'Loading pretrained {backbone}...')
            new_dict = load_pretrain(backbone, self.num_patches, self.conv_proj[0].weight.shape[2], './pretrained')
            self.load_state_dict(new_dict, strict=False)
            logging.info(f'Load pretrained {backbone} sucessfully!')
        if self.freeze_vit:
            for k, p in self.named_parameters():
                if "transformer" in k or "cls_token" in k or "conv_proj" in k or "pos_embedding" in k:
                    p.requires_grad = False
                if "scale" in k or "shift" in k:
                    p.requires_grad = True
        self._cfg = _vit_cfg(depth, heads, dim, mlp_dim, dim_head, channels, frames, frame_patch_size, image_height, image_width,
                             patch_height, patch_width, num_patches)
        from ..vit_engine import VitEngine
        self._engine = VitEngine(self, 'ssf', compute_dtype)

    def init_head_weights(self):
        nn.init.xavier_uniform_(self.mlp_head.weight)
        nn.init.zeros_(self.mlp_head.bias)
        logging.info("Initialize head weight successfully!")

    def train(self, mode=True):
        """Reference quirk preserved (ssf.py:216-228)."""
        if mode:
            super().train(mode)
            if self.freeze_vit:
                self.transformer.eval()
                self.conv_proj.eval()
                self.dropout.eval()
                self.mlp_head.train()
        else:
            for module in self.children():
                module.eval()

    def set_compute_dtype(self, compute_dtype):
        self._engine.set_compute_dtype(compute_dtype)

    def forward(self, img):
        return self._engine(img)
