"""Drop-in for the reference ``src/model/vpt.py`` (``--method shallow_vpt | deep_vpt``): prompt tokens projected by
``prompt_proj`` and inserted after the cls token (``vpt.py:124-161``).  Deep prompts are re-inserted at every layer with the
reference's ``x[:, 1 + prompt_dim:]`` slice (``vpt.py:151-153``), so the sequence length changes per layer."""
import logging

import torch
from torch import nn

from ..utils.load_pretrained import mapping_vit
from .vision_transformer import VisionTransformer


class PromptedVisionTransformer(nn.Module):
    def __init__(self, image_size, image_patch_size, frames, frame_patch_size, dropout: 0.0, emb_dropout: 0.0, num_classes=5, channels=3,
                 dim_head=64, freeze_vit=True, pool='cls', backbone=None, prompt_dropout=0.0, prompt_dim=64, num_prompts=8, deep_prompt=True,
                 compute_dtype=None, **kwargs):
        super().__init__()
        num_layers, num_heads, hidden_dim, mlp_dim = mapping_vit(backbone)
        self.image_size = image_size
        self.num_layers = num_layers
        self.image_patch_size = image_patch_size
        self.hidden_dim = hidden_dim
        self.mlp_dim = mlp_dim
        self.emb_dropout = emb_dropout
        self.dropout = dropout
        self.num_classes = num_classes
        self.deep_prompt = deep_prompt
        with open('deep_prompt.txt', 'a') as f:          # constructor side effect of the reference (vpt.py:54-55)
            f.write(f'Deep prompt: {self.deep_prompt}\n')
        self.prompt_proj = nn.Linear(prompt_dim, hidden_dim)
        self.prompt_dropout = nn.Dropout(prompt_dropout)
        if self.deep_prompt:
            self.deep_prompt_embeddings = nn.Parameter(torch.zeros(num_layers, num_prompts, prompt_dim))
            nn.init.xavier_uniform_(self.deep_prompt_embeddings.data)
        else:
            self.prompt_embeddings = nn.Parameter(torch.zeros(1, num_prompts, prompt_dim))
            nn.init.xavier_uniform_(self.prompt_embeddings.data)
        self.vision_transformer = VisionTransformer(image_size=image_size, image_patch_size=image_patch_size, frames=frames,
                                                    frame_patch_size=frame_patch_size, num_classes=num_classes, dim=hidden_dim, depth=num_layers,
                                                    heads=num_heads, mlp_dim=mlp_dim, pool=pool, channels=channels, dim_head=dim_head,
                                                    dropout=dropout, emb_dropout=emb_dropout, backbone=backbone, compute_dtype=compute_dtype)
        self.freeze_vit = freeze_vit
        self.init_head_weights()
        self.init_promptproj_weights()
        if self.freeze_vit:
            for k, p in self.vision_transformer.named_parameters():
                if "transformer" in k or "cls_token" in k or "conv_proj" in k or "pos_embedding" in k:
                    p.requires_grad = False
        self._cfg = self.vision_transformer._cfg
        from ..vit_engine import VitEngine
        self._engine = VitEngine(self, 'vpt', compute_dtype)

    def init_head_weights(self):
        nn.init.xavier_uniform_(self.vision_transformer.mlp_head.weight)
        nn.init.zeros_(self.vision_transformer.mlp_head.bias)
        logging.info("Initialize head weight successfully!")

    def init_promptproj_weights(self):
        nn.init.xavier_uniform_(self.prompt_proj.weight)
        nn.init.zeros_(self.prompt_proj.bias)
        logging.info("Initialize prompt projector successfully!")

    def train(self, mode=True):
        """Reference quirk preserved (vpt.py:106-119)."""
        if mode:
            super().train(mode)
            if self.freeze_vit:
                self.vision_transformer.transformer.eval()
                self.vision_transformer.conv_proj.eval()
                self.vision_transformer.dropout.eval()
                self.vision_transformer.mlp_head.train()
                self.prompt_proj.train()
        else:
            for module in self.children():
                module.eval()

    def set_compute_dtype(self, compute_dtype):
        self._engine.set_compute_dtype(compute_dtype)

    def forward(self, x: torch.Tensor):
        return self._engine(x)
