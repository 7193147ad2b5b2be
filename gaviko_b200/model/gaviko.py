"""Drop-in replacement for the reference ``src/model/gaviko.py`` (``Gaviko`` and its sub-modules).

Same constructor kwargs, same ``forward(img) -> logits``, same ``named_parameters()`` / ``state_dict()`` names and
shapes (= the trainable-only checkpoint layout of reference ``src/train.py:161-167,478-483``), same freeze rule and the
same ``train()/eval()`` quirks — but forward/backward run the hand-written sm_100a kernels of ``libgvk_sm100a.so``
through ``gaviko_b200.engine``.  The sub-modules below are parameter containers only: their own ``forward`` is never
used and there is no eager / CPU fallback.

Reference anchors: ctor ``model/gaviko.py:328-443``, freeze rule ``:428-434``, ``init_weights`` ``:445-511``,
``train`` ``:513-528``, ``forward`` ``:531-552``.
"""
import logging
import math

import torch
from torch import nn

from ..engine import GavikoEngine
from ..utils.load_pretrained import load_pretrain, mapping_vit
from . import vision_transformer


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


class QuickGELU(nn.Module):
    """x * sigmoid(1.702 x) (reference model/gaviko.py:15-17); evaluated inside the fused row kernels."""

    def forward(self, x):  # pragma: no cover - container only
        raise RuntimeError('gaviko_b200 sub-modules are parameter containers; call the top-level model')


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError('gaviko_b200 sub-modules are parameter containers; call the top-level model')


class PromptRelevantEstimator(_Container):
    """LN(r) -> Linear(r,64) -> GELU -> Linear(64,P) -> Sigmoid (reference model/gaviko.py:20-47)."""

    def __init__(self, latent_dim, num_prompts):
        super().__init__()
        self.cls_analyzer_ = nn.Sequential(nn.LayerNorm(latent_dim), nn.Linear(latent_dim, 64), nn.GELU(),
                                           nn.Linear(64, num_prompts), nn.Sigmoid())

    @property
    def cls_analyzer(self):
        return self.cls_analyzer_

    def __getitem__(self, index):
        return self.cls_analyzer_[index]


class PromptContextFusion(_Container):
    """LN(r) -> Linear(r,1) -> Sigmoid (reference model/gaviko.py:48-70)."""

    def __init__(self, latent_dim):
        super().__init__()
        self.gl_balancer_ = nn.Sequential(nn.LayerNorm(latent_dim), nn.Linear(latent_dim, 1), nn.Sigmoid())

    @property
    def gl_balancer(self):
        return self.gl_balancer_

    def __getitem__(self, index):
        return self.gl_balancer_[index]


class GlobalAttention(_Container):
    def __init__(self, latent_dim, num_prompts):
        super().__init__()
        self.latent_dim = latent_dim
        self.scale = latent_dim ** -0.5
        self.num_prompts = num_prompts
        self.query_proj = nn.Linear(latent_dim, latent_dim)


class LocalAttention(_Container):
    def __init__(self, latent_dim):
        super().__init__()
        self.latent_dim = latent_dim
        self.scale = latent_dim ** -0.5
        self.query_proj = nn.Linear(latent_dim, latent_dim)


class Awakening_Prompt(_Container):
    """Parameters of the gated prompt fusion block (reference model/gaviko.py:121-147)."""

    def __init__(self, dim, num_prompts, prompt_latent_dim=20):
        super().__init__()
        self.latent_dim = prompt_latent_dim
        self.num_prompts = num_prompts
        self.scale = dim ** -0.5
        self.proj_down = nn.Sequential(nn.Linear(dim, self.latent_dim), QuickGELU())
        self.proj_up = nn.Linear(self.latent_dim, dim)
        self.cls_analyzer = PromptRelevantEstimator(self.latent_dim, self.num_prompts)
        self.gl_balancer = PromptContextFusion(self.latent_dim)
        self.global_attention = GlobalAttention(self.latent_dim, self.num_prompts)
        self.local_attention = LocalAttention(self.latent_dim)
        # alias attributes: extra state_dict keys sharing storage, de-duplicated by named_parameters()
        self.global_query = self.global_attention.query_proj
        self.local_query = self.local_attention.query_proj
        self.attend = nn.Softmax(dim=-1)


class LocalSelfAttention(_Container):
    """Parameters of the window-sparse local attention (reference model/gaviko.py:189-227).

    The reference materialises an (N, N) additive {0,-inf} mask; the kernels use its closed form instead
    (allowed j: i_ax - k_ax//2 <= j_ax <= i_ax + k_ax - 1 - k_ax//2), so no ``mask`` attribute exists here.
    """

    def __init__(self, dim, local_k=(3, 6, 6), DHW=None, attn_drop=0.0, proj_drop=0.0, local_dim=20, qkv_bias=False, dtype=torch.float32):
        super().__init__()
        self.dim = dim
        self.scale = dim ** -0.5
        self.latent_dim = local_dim
        self.norm = nn.LayerNorm(dim)
        self.proj_down = nn.Linear(dim, self.latent_dim)
        self.qkv = nn.Linear(self.latent_dim, self.latent_dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj_up = nn.Linear(self.latent_dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.DHW = DHW
        self.local_k = tuple(local_k)


class Transformer(_Container):
    """Parameter layout of reference model/gaviko.py:246-289 (local_attns / prompt_projs shared every `share_factor` layers)."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, num_prompts, prompt_latent_dim, DHW, local_k, share_factor=1,
                 attn_drop=0., proj_drop=0., local_dim=20, dropout=0., dtype=torch.float32):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.num_prompts = num_prompts
        self.depth = depth
        self.share_factor = share_factor
        unique = math.ceil(depth / share_factor)
        self.local_attns = nn.ModuleList([
            LocalSelfAttention(dim, local_k, DHW, attn_drop, proj_drop, local_dim, qkv_bias=False, dtype=dtype) for _ in range(unique)])
        self.prompt_projs = nn.ModuleList([Awakening_Prompt(dim, num_prompts, prompt_latent_dim) for _ in range(unique)])
        self.attns = nn.ModuleList([vision_transformer.Attention(dim, heads, dim_head, dropout) for _ in range(depth)])
        self.mlps = nn.ModuleList([vision_transformer.FeedForward(dim, mlp_dim, dropout) for _ in range(depth)])


class AdaptiveFusionHead(_Container):
    """Linear head over the mean of the P prompt rows and the cls row (reference model/gaviko.py:308-325)."""

    def __init__(self, dim, num_prompts, num_classes):
        super().__init__()
        self.head = nn.Linear(dim, num_classes)
        self.num_prompts = num_prompts


class Gaviko(nn.Module):
    def __init__(self, *, image_size, image_patch_size, frames, frame_patch_size, num_classes, pool='cls', channels=1,
                 dim_head=64, dropout=0., emb_dropout=0., backbone=None, num_prompts=8, prompt_latent_dim=20, local_dim=20,
                 local_k=(3, 6, 6), DHW=(10, 10, 10), attn_drop=0.2, proj_drop=0.2, freeze_vit=False, share_factor=1,
                 fp16=False, compute_dtype=None, **kwargs):
        super().__init__()
        self.dtype = torch.float32 if not fp16 else torch.float16
        print("GAViKO Model Initialization")
        print(f"Using dtype: {self.dtype}")
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
        self.num_prompts = num_prompts
        self.local_dim = local_dim
        self.local_k = local_k
        self.prompt_latent_dim = prompt_latent_dim
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'

        self.conv_proj = nn.Sequential(nn.Conv3d(channels, dim, kernel_size=(frame_patch_size, image_patch_size, image_patch_size),
                                                 stride=(frame_patch_size, image_patch_size, image_patch_size)))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, num_prompts, prompt_latent_dim, DHW, local_k, share_factor,
                                       attn_drop, proj_drop, local_dim, dropout, dtype=self.dtype)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = AdaptiveFusionHead(dim, num_prompts, num_classes)
        scale = dim ** -0.5
        self.prompt_positional_embedding = nn.Parameter(scale * torch.randn(1, num_prompts, dim))
        self.prompt_embeddings = nn.Parameter(torch.randn(1, num_prompts, dim))

        self.freeze_vit = freeze_vit
        if self.freeze_vit:
            for k, p in self.named_parameters():
                if "transformer" in k or "cls_token" in k or "conv_proj" in k or "pos_embedding" in k:
                    p.requires_grad = False
                if "head" in k or "prompt" in k or "local_attn" in k:
                    p.requires_grad = True

        if backbone is not None:
            logging.info(f'Loading pretrained {backbone}...')
            new_dict = load_pretrain(backbone, self.num_patches, self.conv_proj[0].weight.shape[2], './pretrained')
            self.load_state_dict(new_dict, strict=False)
            logging.info(f'Load pretrained {backbone} sucessfully!')
        self.init_weights()

        # engine configuration (not part of the reference surface; optional kwarg `compute_dtype`: 'fp32' | 'bf16')
        self._cfg = dict(depth=depth, heads=heads, dim=dim, mlp_dim=mlp_dim, dim_head=dim_head, channels=channels,
                         grid=(frames // frame_patch_size, image_height // patch_height, image_width // patch_width),
                         fp=frame_patch_size, ps=patch_height, num_prompts=num_prompts, num_patches=num_patches,
                         local_k=tuple(int(k) for k in local_k), DHW=tuple(int(d) for d in DHW), share_factor=share_factor,
                         attn_drop=float(attn_drop), proj_drop=float(proj_drop), local_dim=local_dim, prompt_latent_dim=prompt_latent_dim)
        assert patch_height == patch_width, 'square in-plane patches only'
        assert self._cfg['grid'] == self._cfg['DHW'], 'DHW must equal the patch grid'
        self._engine = GavikoEngine(self, compute_dtype)

    # ------------------------------------------------------------------------------------------
    def init_weights(self, scale_factor=1.0):
        """Initialisation recipe of reference model/gaviko.py:445-511, replayed in the same RNG order (so a seeded
        construction gives the reference's weights): clipped-normal prompts, Xavier side paths, orthogonal queries."""
        sf = scale_factor
        with torch.no_grad():
            self.prompt_embeddings.normal_(0.0, 0.02 * sf).clamp_(-0.04 * sf, 0.04 * sf)
            self.prompt_positional_embedding.normal_(0.0, 0.01 * sf)

        def apply(plan):
            for tensor, kind, arg in plan:
                if tensor is None:
                    continue
                if kind == 'xavier':
                    nn.init.xavier_uniform_(tensor, gain=arg)
                elif kind == 'orth':
                    nn.init.orthogonal_(tensor, gain=arg)
                else:
                    nn.init.constant_(tensor, arg)

        for pp in self.transformer.prompt_projs:
            est, bal = pp.cls_analyzer, pp.gl_balancer
            apply([(pp.proj_down[0].weight, 'xavier', 0.7 * sf), (pp.proj_down[0].bias, 'const', 0.0),
                   (pp.proj_up.weight, 'xavier', 0.7 * sf), (pp.proj_up.bias, 'const', 0.0),
                   (pp.global_query.weight, 'orth', sf), (pp.global_query.bias, 'const', 0.0),
                   (pp.local_query.weight, 'orth', sf), (pp.local_query.bias, 'const', 0.0),
                   (est[1].weight, 'xavier', 1.0), (est[1].bias, 'const', 0.0),
                   (est[3].weight, 'xavier', 1.0), (est[3].bias, 'const', 0.0),       # sigmoid(0) = 0.5 importance
                   (bal[1].weight, 'xavier', 1.0), (bal[1].bias, 'const', 0.5)])      # starts ~0.62 global / 0.38 local
        for la in self.transformer.local_attns:
            # NB: like the reference, proj_up.bias keeps its nn.Linear default (the reference zeroes proj_down.bias twice).
            apply([(la.proj_down.weight, 'xavier', 0.5 * sf), (la.proj_down.bias, 'const', 0.0),
                   (la.qkv.weight, 'xavier', 1.0), (la.qkv.bias, 'const', 0.0),
                   (la.proj_up.weight, 'xavier', 0.5 * sf)])
        apply([(self.mlp_head.head.weight, 'xavier', 1.0), (self.mlp_head.head.bias, 'const', 0.0)])

    def train(self, mode=True):
        """Reference quirk preserved (model/gaviko.py:513-528): returns None; train(False) never clears self.training."""
        if mode:
            super().train(mode)
            if self.freeze_vit:
                self.transformer.eval()
                self.conv_proj.eval()
                self.dropout.eval()
                self.transformer.local_attns.train()
                self.transformer.prompt_projs.train()
                self.mlp_head.train()
        else:
            for module in self.children():
                module.eval()

    def set_compute_dtype(self, compute_dtype):
        """'fp32' (exact FFMA kernels) or 'bf16' (tcgen05 tensor-core GEMMs / attention, fp32 residual stream and side paths)."""
        self._engine.set_compute_dtype(compute_dtype)

    def forward(self, img):
        return self._engine(img)
