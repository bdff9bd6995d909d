"""Silent-interval bit string -> per-sample mask (TEST INFRASTRUCTURE).

Restates M2/tools.py:340-362 == M2/predict.py:232-252 == M1/tools.py:770-792
(and the inline copy M2/dataset.py:171-189) plus the gating multiplies
M2/predict.py:317, M2/dataset.py:193,229.
"""
from itertools import groupby
import numpy as np


def bits_to_sample_mask(length, ratio, bits):
    """Loop form, as in the reference.  bits: iterable of '0'(silent)/'1'."""
    mask = np.zeros(length, dtype=np.float32)
    for i, bit in enumerate(bits):
        lo, hi = int(i * ratio), int((i + 1) * ratio - 1)
        mask[lo:hi] = 1.0 if bit == "0" else 0.0
    pos = 0
    for k, g in groupby(mask.copy()):
        n = len(list(g))
        if n < 5:
            mask[pos:pos + n] = 1 - k
        pos += n
    return mask


def gate_noise(mixed, ratio, bits):
    """noise_sig = mixed_sig * mask  (M2/predict.py:317, M2/dataset.py:229)."""
    return mixed * bits_to_sample_mask(len(mixed), ratio, bits)


def gate_clean(audio, ratio, bits):
    """audio * (1 - mask)  (M2/dataset.py:193)."""
    return audio * (1 - bits_to_sample_mask(len(audio), ratio, bits))


def logits_to_bits(logits, threshold=0.5):
    """sigmoid(logit) >= 0.5 -> '1' non-silent else '0'  (M1/predict.py:30,117-119)."""
    p = 1.0 / (1.0 + np.exp(-np.asarray(logits, dtype=np.float64)))
    return ["".join("1" if v >= threshold else "0" for v in row) for row in np.atleast_2d(p)]
