"""Import the live reference (``/root/reference/src``) with the two shims SURVEY.md §8c describes.

Only usable in the build container (the GPU box has no ``/root/reference``).  Used by
``oracle/make_golden.py`` and by the in-container ``tests/test_oracle_live_reference.py``
(skipped when the reference is absent).  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

REF_SRC = '/root/reference/src'


def available() -> bool:
    return os.path.isdir(REF_SRC)


def load():
    """Returns a namespace with the reference's model classes and FocalLoss."""
    if not available():
        raise RuntimeError('reference not present at ' + REF_SRC)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    # utils/load_pretrained.py:4 imports timm (not installed, needs network)
    sys.modules.setdefault('timm', types.ModuleType('timm'))
    import utils.load_pretrained as lp
    lp.load_pretrain = lambda *a, **k: {}
    import model.gaviko as g
    import model.vision_transformer as vt
    import model.adaptformer as af
    import model.ssf as ssf
    import model.melo as melo
    import model.vpt as vpt
    import model.dvpt as dvpt
    import model.evp as evp
    evp.device = __import__('torch').device('cpu')       # evp.py:19 picks cuda when present; the goldens are CPU fp32
    for m in (g, vt, af, ssf, dvpt, evp):
        m.load_pretrain = lambda *a, **k: {}       # names bound by `from ... import` (e.g. gaviko.py:7)
    from losses.focal_loss import FocalLoss
    return types.SimpleNamespace(Gaviko=g.Gaviko, VisionTransformer=vt.VisionTransformer, AdaptFormer=af.AdaptFormer,
                                 ScalingShiftingFeatures=ssf.ScalingShiftingFeatures, MeLO=melo.MeLO,
                                 PromptedVisionTransformer=vpt.PromptedVisionTransformer, DynamicVisualPromptTuning=dvpt.DynamicVisualPromptTuning,
                                 ExplicitVisualPrompting=evp.ExplicitVisualPrompting,
                                 FocalLoss=FocalLoss,
                                 gaviko=g, vision_transformer=vt)
