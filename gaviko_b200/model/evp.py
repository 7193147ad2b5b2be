"""Placeholder for the reference ``src/model/evp.py`` (``ExplicitVisualPrompting``, ``--method evp``; SURVEY.md §8 f4, the lowest-ranked "next" row).

The reference's scripts import this module unconditionally (``train.py:9``, ``eval.py:9``, ``inference.py:9``), so the drop-in tree must provide
it; the method itself (FFT high-pass handcrafted prompts + a second patch embedding + per-layer prompt MLPs) has no CUDA path yet and says so
loudly instead of falling back to eager PyTorch."""
from torch import nn


class ExplicitVisualPrompting(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError('gaviko_b200: --method evp (ExplicitVisualPrompting) is not implemented (SURVEY.md §8 f4); '
                                  'the other methods (gaviko, linear, bitfit, adaptformer, melo, ssf, shallow_vpt, deep_vpt, dvpt) are')
