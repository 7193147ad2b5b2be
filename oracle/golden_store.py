"""Compact storage of gradients in golden files (TEST INFRASTRUCTURE ONLY).

ViT-B / ViT-L GAViKO have 0.9 M / 2.3 M trainable scalars; two losses of full fp32 gradients would be 7 / 18 MB per case.  The 'subset'
store keeps full gradients of the tensors most likely to expose an error (prompts, head, first / middle / last block) and, for every other
tensor, the sums of consecutive 32-element chunks of the flattened gradient: 1/32 of the size, and for errors without a fixed sign the relative
L2 error of the chunk-sum vector is of the same order as that of the full tensor.
"""
import re

import numpy as np

CHUNK = 32


def stored_in_full(name: str, depth: int) -> bool:
    if name.startswith('prompt_generator.'):      # EVP: biases and three of the per-layer generators in full, the large matrices as chunk sums
        m = re.search(r'lightweight_mlp_(\d+)\.', name)
        if m is not None:
            return int(m.group(1)) in (0, depth // 2, depth - 1)
        return name.endswith('.bias')
    m = re.search(r'\.(\d+)\.', name)
    if m is None:
        return True                      # prompts, head
    return int(m.group(1)) in (0, depth // 2, depth - 1)


def chunk_sums(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    pad = (-a.size) % CHUNK
    if pad:
        a = np.concatenate([a, np.zeros(pad)])
    return a.reshape(-1, CHUNK).sum(1)


def fingerprint(sd) -> dict:
    """{name: (sum, sum of squares)} in float64 over every floating tensor of a state dict."""
    out = {}
    for k, v in sd.items():
        if v.is_floating_point():
            d = v.detach().double()
            out[k] = (float(d.sum()), float((d * d).sum()))
    return out
