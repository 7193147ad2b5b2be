"""Architecture table and pretrained-weight hooks (reference ``src/utils/load_pretrained.py``).

Only ``mapping_vit`` is on the hot path.  The reference's ``load_pretrain`` downloads timm ImageNet-21k weights
(``load_pretrained.py:8-99``); this build has no network and the north star specifies random-init ViT weights, so
``load_pretrain`` keeps its signature and returns an empty dict unless a local ``{save_dir}/{timm_name}`` state_dict
file exists, in which case it is converted with the same key / shape rules (2D->3D kernel inflation, trilinear position
interpolation).
"""
import logging
import os

import torch
import torch.nn.functional as F

_WARNED = False      # warn once per process
_ARCH = {
    'vit-t16': (12, 3, 192, 768),
    'vit-s16': (12, 6, 384, 1536),
    'vit-b16': (12, 12, 768, 3072),
    'vit-l16': (24, 16, 1024, 4096),
}
_TIMM_NAMES = {'vit-b16': 'vit_base_patch16_224_in21k', 'vit-t16': 'vit_tiny_patch16_224_in21k',
               'vit-s16': 'vit_small_patch16_224_in21k', 'vit-l16': 'vit_large_patch16_224_in21k'}


def mapping_vit(backbone):
    """backbone name -> (depth, heads, dim, mlp_dim); ValueError on None / unknown (reference :103-120)."""
    if backbone is None:
        raise ValueError("Backbone must be specified.")
    key = backbone.lower()
    if key not in _ARCH:
        raise ValueError(f"Unsupported backbone: {backbone}. Supported backbones are: {list(_ARCH.keys())}")
    return _ARCH[key]


def _convert(timm_sd, num_patches, depth_dim):
    """timm ViT state_dict -> GAViKO names (reference :56-98)."""
    out = {}
    for key, value in timm_sd.items():
        if key == 'cls_token':
            out[key] = value
        elif key == 'pos_embed':
            cls_pos, grid = value[:, :1], value[:, 1:]
            side = int(grid.shape[1] ** 0.5)
            grid = grid.reshape(1, side, side, -1).permute(0, 3, 1, 2).unsqueeze(2)
            new = round(num_patches ** (1 / 3))
            grid = F.interpolate(grid, size=(new, new, new), mode='trilinear', align_corners=False)
            out['pos_embedding'] = torch.cat([cls_pos, grid.permute(0, 2, 3, 4, 1).reshape(1, new ** 3, -1)], dim=1)
        elif key == 'patch_embed.proj.weight':
            out['conv_proj.0.weight'] = value.mean(dim=1, keepdim=True).unsqueeze(2).repeat(1, 1, depth_dim, 1, 1)
        elif key == 'patch_embed.proj.bias':
            out['conv_proj.0.bias'] = value
        elif key in ('norm.weight', 'norm.bias'):
            out['transformer.' + key] = value
        elif key.startswith('blocks.'):
            for src, dst, group in (('norm1', 'norm', 'attns'), ('attn.qkv', 'to_qkv', 'attns'), ('attn.proj', 'to_out.0', 'attns'),
                                    ('norm2', 'net.0', 'mlps'), ('mlp.fc1', 'net.1', 'mlps'), ('mlp.fc2', 'net.4', 'mlps')):
                if src in key:
                    out[key.replace(src, dst).replace('blocks', f'transformer.{group}')] = value
                    break
    return out


def load_pretrain(backbone, num_patches, depth_dim, save_dir):
    name = _TIMM_NAMES.get(backbone.replace('_', '-').lower())
    path = os.path.join(save_dir, name) if name else None
    if path and os.path.exists(path):
        logging.info(f'Converting local pretrained weights {path}')
        return _convert(torch.load(path, map_location='cpu'), num_patches, depth_dim)
    # The reference downloads the timm ImageNet-21k weights here (load_pretrained.py:24-31).  Offline there is nothing to download: say so loudly,
    # because fine-tuning PEFT modules on a frozen RANDOM backbone is rarely what a user of the reference configs wants.
    global _WARNED
    if _WARNED:
        return {}
    _WARNED = True
    logging.warning(f'gaviko_b200: no local pretrained weights at {path!r} (no network in this build): the frozen ViT backbone keeps its RANDOM '
                    'initialisation.  Place a timm state_dict there (torch.save(timm_model.state_dict(), path)) to fine-tune a pretrained backbone.')
    return {}


def load_vanilla_pretrain(backbone, config):
    m = config['model']
    depth_dim = m['frame_patch_size']
    ih = iw = m['image_size']
    num_patches = (ih // m['image_patch_size']) * (iw // m['image_patch_size']) * (m['frames'] // m['frame_patch_size'])
    return load_pretrain(backbone, num_patches, depth_dim, save_dir='./pretrained')


def load_vanilla_pretrain_with_adapters(backbone, config, checkpoint_path):
    """{**vanilla, **checkpoint} exactly as reference :150-156 (the checkpoint holds the trainable tensors only)."""
    merged = dict(load_vanilla_pretrain(backbone, config))
    merged.update(torch.load(checkpoint_path, map_location='cpu'))
    return merged
