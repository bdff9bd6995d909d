"""Deterministic synthetic noisy-speech clips (TEST INFRASTRUCTURE).

SURVEY.md section 8d recipe.  Mixing follows the reference's add_signals
(M2/tools.py:217-276) with norm=0.5 and the SNR list of M2/dataset.py:34;
speech is forced silent on the ground-truth silent intervals exactly like
M2/dataset.py:193 (audio * (1 - mask)); noise_sig = mixed * mask
(M2/dataset.py:229).
"""
import numpy as np
from .gating import bits_to_sample_mask

SNRS = [-10, -7, -3, 0, 3, 7, 10]        # M2/dataset.py:34
FPS = 30.0


def power_of_signal(x):                   # M2/tools.py:213-214
    return np.sum(np.abs(x ** 2))


def add_signals(signal, noise, snr, norm=0.5):
    """M2/tools.py:217-276 for a single noise track."""
    sp = power_of_signal(signal)
    pn = sp / np.power(10, snr / 10)
    ret = np.copy(signal)
    if sp == 0:
        new_noise = noise
    else:
        ratio = np.sqrt(power_of_signal(noise)) / np.sqrt(pn)
        new_noise = noise if ratio == 0 else noise / ratio
    ret = ret + new_noise
    if norm:
        scale = np.max(np.abs(ret)) / norm
        if scale != 0:
            return ret / scale, signal / scale, new_noise / scale
    return ret, signal, new_noise


def make_bits(rng, n_frames):
    bits, state = [], int(rng.integers(0, 2))
    while len(bits) < n_frames:
        run = int(rng.integers(3, 21))
        bits.extend([str(state)] * run)
        state ^= 1
    return "".join(bits[:n_frames])


def make_clip(index, length=32000, sr=16000):
    """Returns dict of float32 waveforms + bit string + label vector."""
    rng = np.random.default_rng(1234 + index)
    n_frames = int(round(length / sr * FPS))
    bits = make_bits(rng, n_frames)
    ratio = sr / FPS
    t = np.arange(length) / sr
    f0 = rng.uniform(100, 250)
    speech = np.zeros(length)
    for h in range(1, 6):
        speech += np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi)) / h
    speech *= 0.5 * (1 - np.cos(2 * np.pi * 4.0 * t))
    mask = bits_to_sample_mask(length, ratio, bits).astype(np.float64)
    speech = speech * (1 - mask)
    white = rng.standard_normal(length)
    noise = np.zeros(length)
    acc = 0.0
    for i in range(length):                    # 1-pole low-pass, alpha = 0.95
        acc = 0.95 * acc + 0.05 * white[i]
        noise[i] = acc
    snr = SNRS[index % len(SNRS)]
    mixed, clean, full_noise = add_signals(speech, noise, snr, norm=0.5)
    return {
        "mixed": mixed.astype(np.float32),
        "clean": clean.astype(np.float32),
        "full_noise": full_noise.astype(np.float32),
        "noise": (mixed * mask).astype(np.float32),
        "mask": mask.astype(np.float32),
        "bits": bits,
        "label": np.array([float(b) for b in bits], dtype=np.float32),
        "snr": snr,
    }


def make_batch(batch, length=32000, sr=16000, start=0):
    clips = [make_clip(start + i, length, sr) for i in range(batch)]
    out = {k: np.stack([c[k] for c in clips]) for k in
           ("mixed", "clean", "full_noise", "noise", "mask", "label")}
    out["bits"] = [c["bits"] for c in clips]
    return out
