"""Device-side input transform in front of the hot path.

The reference rescales every volume to [0, 1] on the CPU, inside the DataLoader workers, as the last transform of each pipeline
(``tio.RescaleIntensity(out_min_max=(0,1))`` at train.py:53,57,61; eval.py:31; inference.py:30).  `RescaleIntensity` below is the same
transform for a batch that is already on the GPU: load the raw volumes, copy them to the device, then call it — the fp32 result is
bit-identical to torchio's (same order of fp32 operations), and `out_dtype=torch.bfloat16` writes the model's compute dtype directly."""
import torch

from . import ops


class RescaleIntensity:
    """torchio.RescaleIntensity(out_min_max=..., percentiles=(0, 100)) on CUDA tensors; every x[b] is rescaled on its own."""

    def __init__(self, out_min_max=(0, 1), percentiles=(0, 100), out_dtype=None):
        if tuple(percentiles) != (0, 100):
            raise NotImplementedError('RescaleIntensity: only the default percentiles (0, 100) are implemented (the reference uses the default)')
        self.out_min, self.out_max = float(out_min_max[0]), float(out_min_max[1])
        if self.out_min > self.out_max:
            raise ValueError(f'out_min_max must be increasing, got {out_min_max}')      # torchio's own check
        self.out_dtype = out_dtype

    def __call__(self, x, inplace=False):
        single = x.dim() == 4                      # (C, D, H, W): one volume, as the Dataset sees it
        xb = x.unsqueeze(0) if single else x
        xb = xb.float().contiguous()
        out = ops.rescale_intensity(xb, self.out_min, self.out_max, out=xb if inplace and self.out_dtype in (None, torch.float32) else None,
                                    out_dtype=self.out_dtype)
        return out[0] if single else out
