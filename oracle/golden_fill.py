"""Deterministic, reference-independent weight fill used to pin parity.

The reference's random init cannot travel to the GPU box (``/root/reference``
does not exist there) and a ViT state_dict is too large to commit, so golden
vectors are generated with every tensor of ``state_dict()`` overwritten by the
procedure below, which depends only on the tensor's *name* and *shape*.  A
drop-in module with the same names/shapes (the checkpoint contract,
reference ``src/train.py:161-167,478-483``) therefore reproduces the exact same weights.
"""
import zlib

import torch


def _scale_for(name: str, shape) -> tuple:
    """(mean, std) for one named tensor; magnitudes chosen so activations stay O(1)."""
    nd = len(shape)
    leaf = name.rsplit('.', 1)[-1]
    if 'ssf_scale' in leaf:
        return 1.0, 0.1
    if 'ssf_shift' in leaf:
        return 0.0, 0.05
    if name.endswith('pos_embedding'):
        return 0.0, 0.5
    if name.endswith('cls_token'):
        return 0.0, 0.5
    if name.endswith('prompt_positional_embedding'):
        return 0.0, 0.1
    if 'prompt_embeddings' in name:
        return 0.0, 0.5
    if nd == 1:
        if leaf == 'weight':          # LayerNorm gain
            return 1.0, 0.1
        return 0.0, 0.05              # every bias
    fan_in = 1
    for s in shape[1:]:
        fan_in *= int(s)
    return 0.0, fan_in ** -0.5        # Linear / Conv3d weights


@torch.no_grad()
def golden_fill(module_or_sd, seed: int = 0):
    """Overwrite every floating tensor in place; returns the state_dict used.

    Alias keys that share storage (reference ``model/gaviko.py:144-145``) are
    filled once, under the lexicographically first name.
    """
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, 'state_dict') else module_or_sd
    seen = set()
    for name in sorted(sd.keys()):
        t = sd[name]
        if not torch.is_floating_point(t):
            continue
        key = (t.data_ptr(), tuple(t.shape))
        if key in seen:
            continue
        seen.add(key)
        g = torch.Generator(device='cpu')
        g.manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
        mean, std = _scale_for(name, t.shape)
        vals = torch.randn(t.shape, generator=g, dtype=torch.float32) * std + mean
        t.copy_(vals.to(t.dtype))
    return sd


def golden_volume(batch: int, frames: int, height: int, width: int, seed: int = 1234, channels: int = 1):
    """Synthetic volumes in [0,1) (post-RescaleIntensity range, reference ``src/train.py:53-57``)."""
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    return torch.rand(batch, channels, frames, height, width, generator=g, dtype=torch.float32)


def golden_labels(batch: int, num_classes: int = 5, seed: int = 1234):
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    return torch.randint(0, num_classes, (batch,), generator=g, dtype=torch.int64)


def golden_eval_volume(seed: int, frames: int, height: int, width: int):
    """One structured synthetic volume for the argmax-identity eval set: a smooth random field (trilinear upsampling of a coarse
    6x8x8 lattice) with per-volume contrast, offset and noise level, clipped to [0, 1].  Unlike ``golden_volume`` (i.i.d. uniform
    voxels, which all look alike to the network) these spread the logits over several classes."""
    import torch.nn.functional as F
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    coarse = torch.randn(1, 1, 6, 8, 8, generator=g)
    field = F.interpolate(coarse, size=(frames, height, width), mode='trilinear', align_corners=True)
    contrast, offset, noise = torch.rand(3, generator=g).tolist()
    v = 0.5 + (0.15 + 0.6 * contrast) * field + (offset - 0.5) * 0.6 + (0.02 + 0.2 * noise) * torch.randn(field.shape, generator=g)
    return v.clamp_(0.0, 1.0)
