"""TEST INFRASTRUCTURE ONLY — CPU restatement of torchio.RescaleIntensity for SURVEY.md §8 f2 (never imported by the product path).

The algorithm lives in a third-party dependency that is absent from /root/reference and from this image: torchio==0.20.16
(requirements.txt:6), class RescaleIntensity, method `rescale` (torchio/transforms/preprocessing/intensity/rescale_intensity.py).  Its
published algorithm for the reference's call sites — `tio.RescaleIntensity(out_min_max=(0,1))`, default percentiles (0, 100), no mask —
at train.py:53,57,61, eval.py:31 and inference.py:30 is restated below in numpy, operation for operation:

    array = tensor.clone().float().numpy()
    cutoff = np.percentile(array, (0, 100)); np.clip(array, *cutoff, out=array)       # no-op for (0, 100)
    in_min, in_max = array.min(), array.max()
    array -= in_min; in_range = in_max - in_min
    if in_range == 0: warn and return the input unchanged
    array /= in_range; array *= out_max - out_min; array += out_min

Parity unpinned: the reference holds no fixture for this transform and torchio cannot be imported here; the oracle is anchored by
known-answer cases (tests/test_rescale.py) derived by hand from the formula above."""
import numpy as np


def rescale_intensity(volume, out_min=0.0, out_max=1.0):
    """volume: ndarray of any shape (one image, all channels together, as torchio treats it) -> float32 ndarray."""
    array = np.array(volume, dtype=np.float32, copy=True)
    lo, hi = np.percentile(array, (0, 100))
    np.clip(array, lo, hi, out=array)
    in_min, in_max = array.min(), array.max()
    array -= in_min
    in_range = in_max - in_min
    if in_range == 0:
        return np.array(volume, dtype=np.float32, copy=True)
    array /= in_range
    array *= np.float32(out_max - out_min)
    array += np.float32(out_min)
    return array
