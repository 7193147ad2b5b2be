"""Drop-in replacement for the reference ``src/model/evp.py`` (``ExplicitVisualPrompting``, ``--method evp``).

Same constructor kwargs, ``forward(img) -> logits``, parameter names / shapes / freeze rule (``model/evp.py:293-298``), ``train()`` quirk
(``:305-320``) and RNG-order-identical construction (nn defaults, then ``PromptGenerator.apply(_init_weights)``: truncated-normal Linear
weights, zero biases, ``:54-68``).  Forward / backward run the sm_100a kernels through ``gaviko_b200.vit_engine.VitEngine`` (kind 'evp'):
the frozen ViT blocks, plus per layer ``x[:, 1:] += shared_mlp(GELU(lightweight_mlp_i(handcrafted + embedding)))`` where
``embedding = embedding_generator(conv_proj(img))`` and ``handcrafted`` is a second patch embedding of the high-passed volume
(``:72-95,124-146``).  Sub-modules are parameter containers.
"""
import logging
import math

import torch
from torch import nn

from ..utils.load_pretrained import load_pretrain, mapping_vit
from .vision_transformer import Transformer, _Container, _vit_cfg, pair


class PatchEmbed(_Container):
    """``proj`` = the Conv3d patch embedding (reference model/evp.py:148-163); evaluated as a gathered-patch GEMM by the engine."""

    def __init__(self, img_size=160, frames=120, image_patch_size=16, frame_patch_size=12, in_chans=3, dim=768):
        super().__init__()
        self.img_size = img_size
        self.proj = nn.Conv3d(in_chans, dim, kernel_size=(frame_patch_size, image_patch_size, image_patch_size),
                              stride=(frame_patch_size, image_patch_size, image_patch_size))


def _trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """The reference's private truncated normal (model/evp.py:165-206): uniform in the CDF interval -> erfinv -> scale -> clamp.  Restated
    because the RNG consumption (one ``uniform_`` per tensor) is part of the seeded-construction contract."""
    def norm_cdf(x):
        return (1. + math.erf(x / math.sqrt(2.))) / 2.
    with torch.no_grad():
        lo, hi = norm_cdf((a - mean) / std), norm_cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * hi - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
    return tensor


class PromptGenerator(_Container):
    """Parameters of the prompt generator (reference model/evp.py:24-95): ``shared_mlp`` r -> dim, ``embedding_generator`` dim -> r,
    ``lightweight_mlp_{i}`` = [Linear(r, r), GELU] per layer, ``prompt_generator`` = PatchEmbed(in -> r); r = dim // scale_factor."""

    def __init__(self, scale_factor, dim, depth, input_type, freq_nums, handcrafted_tune, embedding_tune, img_size, frames, image_patch_size,
                 frame_patch_size, channels):
        super().__init__()
        self.mode = 'stack'
        self.scale_factor = scale_factor
        self.embed_dim = dim
        self.input_type = input_type
        self.freq_nums = freq_nums
        self.depth = depth
        self.handcrafted_tune = handcrafted_tune
        self.embedding_tune = embedding_tune
        r = self.embed_dim // self.scale_factor
        self.shared_mlp = nn.Linear(r, self.embed_dim)
        self.embedding_generator = nn.Linear(self.embed_dim, r)
        for i in range(self.depth):
            setattr(self, 'lightweight_mlp_{}'.format(str(i)), nn.Sequential(nn.Linear(r, r), nn.GELU()))
        self.prompt_generator = PatchEmbed(img_size=img_size, frames=frames, image_patch_size=image_patch_size, frame_patch_size=frame_patch_size,
                                           in_chans=channels, dim=r)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.Conv2d):      # never true here (the patch embedding is a Conv3d): kept because the reference has the branch
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
            m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                m.bias.data.zero_()


class ExplicitVisualPrompting(nn.Module):
    def __init__(self, *, image_size, image_patch_size, frames, frame_patch_size, num_classes, pool='cls', channels=3, dim_head=64,
                 dropout=0., emb_dropout=0., backbone=None, freeze_vit=False, scale_factor=32, input_type='fft', freq_nums=0.25,
                 handcrafted_tune=True, embedding_tune=True, compute_dtype=None, **kwargs):
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
        if input_type != 'fft':
            # the reference's other input types reference attributes it never creates (lap_pyramid, prompt: model/evp.py:98-110)
            raise NotImplementedError("gaviko_b200 EVP implements input_type='fft' (the only one the reference can run)")
        self.conv_proj = PatchEmbed(image_size, frames, image_patch_size, frame_patch_size, channels, dim)
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Linear(dim, num_classes)
        self.scale_factor = scale_factor
        self.input_type = input_type
        self.freq_nums = freq_nums
        self.handcrafted_tune = handcrafted_tune
        self.embedding_tune = embedding_tune
        self.prompt_generator = PromptGenerator(scale_factor, dim, depth, input_type, freq_nums, handcrafted_tune, embedding_tune, image_size, frames,
                                                image_patch_size, frame_patch_size, channels)
        if backbone is not None:
            logging.info(f'Loading pretrained {backbone}...')
            new_dict = load_pretrain(backbone, self.num_patches, self.conv_proj.proj.weight.shape[2], './pretrained')
            self.load_state_dict(new_dict, strict=False)
            logging.info(f'Load pretrained {backbone} sucessfully!')
        self.freeze_vit = freeze_vit
        if self.freeze_vit:
            for k, p in self.named_parameters():
                if "transformer" in k or "cls_token" in k or "conv_proj" in k or "pos_embedding" in k:
                    p.requires_grad = False
                if "prompt_generator" in k:
                    p.requires_grad = True
        self._cfg = _vit_cfg(depth, heads, dim, mlp_dim, dim_head, channels, frames, frame_patch_size, image_height, image_width,
                             patch_height, patch_width, num_patches)
        self._cfg['evp_rank'] = dim // scale_factor
        from ..vit_engine import VitEngine
        self._engine = VitEngine(self, 'evp', compute_dtype)

    def init_head_weights(self):
        nn.init.xavier_uniform_(self.mlp_head.weight)
        nn.init.zeros_(self.mlp_head.bias)
        logging.info("Initialize head weight successfully!")

    def train(self, mode=True):
        """Reference quirk preserved (model/evp.py:305-320): returns None; train(False) never clears self.training."""
        if mode:
            super().train(mode)
            if self.freeze_vit:
                self.transformer.eval()
                self.conv_proj.eval()
                self.dropout.eval()
                self.mlp_head.train()
                self.prompt_generator.train()
        else:
            for module in self.children():
                module.eval()

    def set_compute_dtype(self, compute_dtype):
        self._engine.set_compute_dtype(compute_dtype)

    def forward(self, img):
        return self._engine(img)
