"""Build a drop-in variant the way reference src/train.py:111-153 does for its --method switch."""
from gaviko_b200.model.adaptformer import AdaptFormer
from gaviko_b200.model.dvpt import DynamicVisualPromptTuning
from gaviko_b200.model.evp import ExplicitVisualPrompting
from gaviko_b200.model.melo import MeLO
from gaviko_b200.model.ssf import ScalingShiftingFeatures
from gaviko_b200.model.vision_transformer import VisionTransformer
from gaviko_b200.model.vpt import PromptedVisionTransformer


def build_variant(method, kw):
    if method in ('linear', 'bitfit'):
        m = VisionTransformer(**kw)
        for k, v in m.named_parameters():                       # train.py:114-137
            v.requires_grad = ('head' in k) if method == 'linear' else (('bias' in k) or ('head' in k))
        return m
    if method == 'adaptformer':
        return AdaptFormer(**kw)
    if method == 'ssf':
        return ScalingShiftingFeatures(**kw)
    if method == 'melo':
        return MeLO(vit=VisionTransformer(**kw), **kw)           # train.py:145-147
    if method in ('deep_vpt', 'shallow_vpt'):
        return PromptedVisionTransformer(**kw)
    if method == 'dvpt':
        return DynamicVisualPromptTuning(**kw)
    if method == 'evp':
        return ExplicitVisualPrompting(**kw)
    raise ValueError(method)
