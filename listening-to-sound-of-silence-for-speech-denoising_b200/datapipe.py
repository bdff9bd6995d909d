"""Training items built ON THE DEVICE from waveform crops.

The reference builds every item in CPU DataLoader workers (60-70 processes): bit string -> sample mask, clean = audio * (1 - mask),
add_signals at a random SNR with norm 0.5, noise_sig = mixed * mask, four librosa STFTs and the cRM target
(M2/dataset.py:144-320; M1/dataset.py:215-352).  Here a Dataset only has to deliver the raw crops (clean speech, a noise
interval, the ground-truth bits, an SNR); everything else is six kernel launches for the whole batch, and the item dicts keep the
reference's keys and layouts so the agents' `forward(data)` takes them unchanged.
"""
import random

import torch

from . import ops, tools, transform

SNRS = [-10, -7, -3, 0, 3, 7, 10]          # M2/dataset.py:34


def crop_noise(noises, length, batch, device, rng=random):
    """add_noise_to_audio's random interval of a random noise track (M2/tools.py:279-292, M2/dataset.py:206) for every clip of
    the batch.  noises: list of 1-D CUDA/CPU tensors, each at least `length` long."""
    out = torch.empty(batch, length, device=device, dtype=torch.float32)
    for i in range(batch):
        n = rng.choice(noises)
        if n.numel() < length:
            raise ValueError(f"noise track of {n.numel()} samples is shorter than the {length}-sample clip")
        start = rng.randint(0, n.numel() - length)
        out[i].copy_(n[start:start + length], non_blocking=True)
    return out


def make_joint_items(audio, noise, snr_db, bits, sr=16000, fps=30.0, norm=0.5, want_waves=False):
    """The stage-2 item dict (M2/dataset.py:311-320) for a whole batch.
    audio, noise (B, L) CUDA fp32 crops; snr_db (B,) floats (tensor or list); bits (B, n) uint8, 0 = silent (or bit strings)."""
    ops.init()
    dev = audio.device
    if not torch.is_tensor(bits):
        bits = tools.bits_to_tensor(bits, dev)
    snr = torch.as_tensor(snr_db, dtype=torch.float32, device=dev)
    ratio = sr / fps
    B = audio.shape[0]
    audio = ops.gate_wave(audio.contiguous(), bits.contiguous(), ratio, 2)       # audio * (1 - mask)        M2/dataset.py:193
    mixed_w, clean_w, full_w = ops.add_signals(audio, noise.contiguous(), snr, norm)      # M2/dataset.py:217
    noise_w = ops.gate_wave(mixed_w, bits, ratio, 1)                               # noise_sig = mixed * mask   M2/dataset.py:229
    spec = transform.stft_batch(torch.cat([mixed_w, clean_w, noise_w, full_w]))    # M2/dataset.py:234-237, one launch
    mixed, clean, noise_s, full = spec[:B], spec[B:2 * B], spec[2 * B:3 * B], spec[3 * B:]
    item = {"mixed": mixed, "clean": clean, "noise": noise_s, "full_noise": full,
            "mask": ops.crm_forward(clean, mixed),                                 # fast_cRM_sigmoid           M2/dataset.py:239
            "start": 0, "bitstream": bits}
    if want_waves:
        item["waves"] = {"mixed": mixed_w, "clean": clean_w, "noise": noise_w, "full_noise": full_w}
    return item


def make_sid_items(audio, noise, snr_db, bits, sr=16000, fps=30.0, norm=0.5, clean_audio=True):
    """The stage-1 item dict (M1/dataset.py:348-352): {"label": (B, n) float 1 = non-silent, "audio": (B, 2, 256, T)}."""
    ops.init()
    dev = audio.device
    if not torch.is_tensor(bits):
        bits = tools.bits_to_tensor(bits, dev)
    ratio = sr / fps
    if clean_audio:                                                                # M1/dataset.py:243-262
        audio = ops.gate_wave(audio.contiguous(), bits.contiguous(), ratio, 2)
        audio, _, _ = ops.add_signals(audio, noise.contiguous(), torch.as_tensor(snr_db, dtype=torch.float32, device=dev), norm)
    return {"label": bits.float(), "audio": transform.stft_batch(audio)}
