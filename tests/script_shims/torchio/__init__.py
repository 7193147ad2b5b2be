"""Stand-in for torchio (absent here): the transforms the scripts build (train.py:37-61, eval.py:30-32, inference.py:29-31).  The random
augmentations are identities; RescaleIntensity is the per-volume min-max rescale (percentiles (0, 100)) on a numpy (C, D, H, W) array."""
import numpy as np


class _Identity:
    def __init__(self, *a, **k):
        pass

    def __call__(self, x):
        return x


class RandomAffine(_Identity): pass
class RandomFlip(_Identity): pass
class RandomNoise(_Identity): pass
class RandomBiasField(_Identity): pass
class RandomBlur(_Identity): pass
class RandomMotion(_Identity): pass


class Compose:
    def __init__(self, transforms, p=1):
        self.transforms = list(transforms)

    def __call__(self, x):
        for t in self.transforms:
            x = t(x)
        return x


class OneOf(Compose):
    pass


class RescaleIntensity:
    def __init__(self, out_min_max=(0, 1), percentiles=(0, 100)):
        self.lo, self.hi = out_min_max

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float32)
        mn, mx = x.min(), x.max()
        if mx == mn:
            return x
        return ((x - mn) / (mx - mn) * (self.hi - self.lo) + self.lo).astype(np.float32)
