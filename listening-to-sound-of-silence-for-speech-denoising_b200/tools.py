"""Silent-interval bit strings on the device (M2/tools.py:340-362, M2/predict.py:232-252, M1/predict.py:117-119)."""
import torch

from . import ops


def bits_to_tensor(bit_strings, device):
    """['0101..', ...] ('0' = silent) -> (B, n) uint8."""
    n = len(bit_strings[0])
    assert all(len(b) == n for b in bit_strings), "bit strings of one batch must have equal length"
    return torch.tensor([[1 if c == "1" else 0 for c in b] for b in bit_strings], dtype=torch.uint8, device=device)


def logits_to_bits(logits, threshold=0.5):
    """sigmoid(logit) >= threshold -> 1 (non-silent)  (M1/predict.py:30,117-119)."""
    return (torch.sigmoid(logits) >= threshold).to(torch.uint8)


def convert_bitstreammask_to_audiomask(audio, ratio, bits):
    """Per-sample 0/1 mask (1 = silent) for waveforms (B, L); bits (B, n) uint8 tensor or list of strings."""
    ops.init()
    if not torch.is_tensor(bits):
        bits = bits_to_tensor(bits, audio.device)
    _, mask = ops.gate_wave(audio.contiguous(), bits.contiguous(), ratio, 1, want_mask=True)
    return mask


def gate_noise(mixed, ratio, bits):
    """noise_sig = mixed_sig * mask  (M2/predict.py:317, M2/dataset.py:229)."""
    ops.init()
    if not torch.is_tensor(bits):
        bits = bits_to_tensor(bits, mixed.device)
    return ops.gate_wave(mixed.contiguous(), bits.contiguous(), ratio, 1)
