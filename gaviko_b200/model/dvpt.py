"""Drop-in replacement for the reference ``src/model/dvpt.py`` (``DynamicVisualPromptTuning``, ``--method dvpt``).

Same constructor kwargs, ``forward(img) -> logits``, parameter names / shapes / freeze rule (``model/dvpt.py:84-163``) and ``train()`` quirk
(``:170-184``); forward / backward run the sm_100a kernels through ``gaviko_b200.dvpt_engine``.  Sub-modules are parameter containers.
"""
import logging

import torch
from torch import nn

from ..utils.load_pretrained import load_pretrain, mapping_vit
from .vision_transformer import Attention, FeedForward, _Container, _vit_cfg, pair


class QuickGELU(_Container):
    """x * sigmoid(1.702 x) (reference model/dvpt.py:21-23); evaluated inside the kernels."""


class share_MLP(_Container):
    """Parameters of the prompt side path (reference model/dvpt.py:25-35): 768 -> 20 -> 768 projections and the scalar gate (zero-initialised)."""

    def __init__(self, d_model, num_prompts):
        super().__init__()
        self.latent_dim = 20
        self.prompt_key_proj_d = nn.Linear(d_model, self.latent_dim)
        self.prompt_key_proj_u = nn.Linear(self.latent_dim, d_model)
        self.prompt_gate = torch.nn.Parameter(torch.zeros(1))
        self.gellu = QuickGELU()
        self.softmax = nn.Softmax(dim=-1)
        self.num = num_prompts
        self.scale = d_model ** -0.5


class ResidualAttentionBlock(_Container):
    def __init__(self, dim, heads, dim_head, mlp_dim, num_prompts, dropout):
        super().__init__()
        self.attn = Attention(dim, heads, dim_head, dropout)
        self.mlp = FeedForward(dim, mlp_dim, dropout)
        self.prompt_proj = share_MLP(dim, num_prompts)


class Transformer(_Container):
    """layers.{i}.0 = ResidualAttentionBlock, final norm (reference model/dvpt.py:65-82)."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, num_prompts, dropout=0., pool='cls'):
        super().__init__()
        self.num = num_prompts
        self.norm = nn.LayerNorm(dim)
        self.pool = pool
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([ResidualAttentionBlock(dim, heads, dim_head, mlp_dim, num_prompts, dropout)]))


class DynamicVisualPromptTuning(nn.Module):
    def __init__(self, *, image_size, image_patch_size, frames, frame_patch_size, num_classes, pool='cls', channels=3, dim_head=64,
                 dropout=0., emb_dropout=0., num_prompts=50, freeze_vit=False, backbone=None, compute_dtype=None, **kwargs):
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
        scale = dim ** -0.5
        self.prompt_positional_embedding = nn.Parameter(scale * torch.randn(1, num_prompts, dim))
        self.prompt_embeddings = nn.Parameter(torch.randn(1, num_prompts, dim))
        self.conv_proj = nn.Sequential(nn.Conv3d(channels, dim, kernel_size=(frame_patch_size, image_patch_size, image_patch_size),
                                                 stride=(frame_patch_size, image_patch_size, image_patch_size)))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, num_prompts, dropout, pool)
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
                if "prompt" in k or "head" in k:
                    p.requires_grad = True
        self._cfg = _vit_cfg(depth, heads, dim, mlp_dim, dim_head, channels, frames, frame_patch_size, image_height, image_width,
                             patch_height, patch_width, num_patches)
        self._cfg['num_prompts'] = num_prompts
        from ..dvpt_engine import DvptEngine
        self._engine = DvptEngine(self, compute_dtype)

    def init_head_weights(self):
        nn.init.xavier_uniform_(self.mlp_head.weight)
        nn.init.zeros_(self.mlp_head.bias)
        logging.info("Initialize head weight successfully!")

    def train(self, mode=True):
        """Reference quirk preserved (model/dvpt.py:170-184): returns None; train(False) never clears self.training."""
        if mode:
            super().train(mode)
            if self.freeze_vit:
                self.transformer.eval()
                self.conv_proj.eval()
                self.dropout.eval()
                self.mlp_head.train()
                for layer in self.transformer.layers:
                    layer[0].prompt_proj.train()
        else:
            for module in self.children():
                module.eval()

    def set_compute_dtype(self, compute_dtype):
        self._engine.set_compute_dtype(compute_dtype)

    def forward(self, img):
        return self._engine(img)
