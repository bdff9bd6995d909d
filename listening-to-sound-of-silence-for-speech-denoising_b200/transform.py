"""Signal transforms with the reference's names and argument meaning (M2/transform.py == M1/transform.py), computed by
the sm_100a kernels.  numpy-in / numpy-out functions keep the reference's layouts; the *_batch functions are the
batched device-tensor forms the training / inference pipeline uses.

  fast_stft                 M2/transform.py:188-193   (librosa.stft(data, 510, 158, 400) -> (256, T, 2))
  fast_istft                M2/transform.py:196-202   (librosa.istft -> (158 (T-1),))
  fast_icRM_sigmoid         M2/transform.py:141-153
  batch_fast_icRM_sigmoid   M2/transform.py:156-169   (differentiable w.r.t. the mask)
"""
import numpy as np
import torch

from . import layers as L
from . import ops

N_FFT = 510          # M2/transform.py:6
HOP_LENGTH = 158     # M2/transform.py:7
WIN_LENGTH = 400     # M2/transform.py:8


def _check_fixed(n_fft, hop_length, win_length):
    if (n_fft, hop_length, win_length) != (N_FFT, HOP_LENGTH, WIN_LENGTH):
        raise ValueError("sos_b200 builds the reference's fixed transform only: n_fft=510, hop_length=158, win_length=400")


def _dev():
    ops.init()
    return torch.device("cuda", torch.cuda.current_device())


def stft_batch(waves, bits=None, ratio=None, gate_mode=0, fused_gate=False):
    """waves (B, L) CUDA fp32 -> (B, 2, 256, T).  Optional silent-interval gating of the waveform first:
    gate_mode 1 = waves * mask (noise gate), 2 = waves * (1 - mask); bits (B, n) uint8, 0 = silent."""
    ops.init()
    return ops.stft(waves.contiguous(), bits, ratio, gate_mode, fused_gate)


def istft_batch(spec, crm=None):
    """(B, 2, 256, T) -> (B, 158 (T-1)); with `crm` the cRM recovery of the mixture `spec` is fused in."""
    ops.init()
    return ops.istft(spec.contiguous(), None if crm is None else crm.contiguous())


def stft_chunked(waves, frames_per_chunk=256):
    """Long-form STFT (BASELINE configs[4]: 10 s clips): the frame axis is processed in chunks of `frames_per_chunk`, each from a
    waveform segment that overlaps its neighbours by two hops so that no kept frame touches a segment edge; only the global
    ends are reflect-padded.  Bit-identical to `stft_batch` on the whole waveform."""
    B, L = waves.shape
    T = 1 + L // HOP_LENGTH
    out = torch.empty(B, 2, 256, T, device=waves.device, dtype=torch.float32)
    for t0 in range(0, T, frames_per_chunk):
        t1 = min(T, t0 + frames_per_chunk)
        lead = 2 if t0 > 0 else 0
        a = HOP_LENGTH * (t0 - lead)
        b = L if t1 == T else min(L, HOP_LENGTH * (t1 - 1 + 2) + 1)
        if b < L and b - HOP_LENGTH * (t1 - 1) < N_FFT // 2 + 1:
            b = L
        seg = stft_batch(waves[:, a:b].contiguous())
        out[:, :, :, t0:t1] = seg[:, :, :, lead:lead + (t1 - t0)]
    return out


def istft_chunked(spec, frames_per_chunk=256, crm=None):
    """Streaming overlap-add: the inverse transform of frame chunks (each with a two-frame margin on both sides) written into
    consecutive 158-sample hops of the output.  Bit-identical to `istft_batch` on the whole spectrogram."""
    B, _, F, T = spec.shape
    out = torch.empty(B, HOP_LENGTH * (T - 1), device=spec.device, dtype=torch.float32)
    for t0 in range(0, T - 1, frames_per_chunk):
        t1 = min(T - 1, t0 + frames_per_chunk)                 # output hops [t0, t1)
        ta, tb = max(0, t0 - 2), min(T, t1 + 3)
        seg = istft_batch(spec[:, :, :, ta:tb].contiguous(), None if crm is None else crm[:, :, :, ta:tb].contiguous())
        s0 = HOP_LENGTH * (t0 - ta)
        out[:, HOP_LENGTH * t0:HOP_LENGTH * t1] = seg[:, s0:s0 + HOP_LENGTH * (t1 - t0)]
    return out


def fast_stft(data, power=False, n_fft=N_FFT, hop_length=HOP_LENGTH, win_length=WIN_LENGTH):
    _check_fixed(n_fft, hop_length, win_length)
    if power:
        raise NotImplementedError("power=True is not on the hot path")
    w = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32), device=_dev())[None]
    return stft_batch(w)[0].permute(1, 2, 0).contiguous().cpu().numpy()


def fast_istft(F, power=False, hop_length=HOP_LENGTH, win_length=WIN_LENGTH):
    _check_fixed(N_FFT, hop_length, win_length)
    if power:
        raise NotImplementedError("power=True is not on the hot path")
    s = torch.as_tensor(np.ascontiguousarray(F, dtype=np.float32), device=_dev()).permute(2, 0, 1)[None].contiguous()
    return istft_batch(s)[0].cpu().numpy()


def batch_fast_icRM_sigmoid(Y, crm, a=0.1, b=0):
    ops.init()
    return L.ICRM.apply(Y, crm, float(a), float(b))


def fast_icRM_sigmoid(Y, crm):
    dev = _dev()
    Yt = torch.as_tensor(np.ascontiguousarray(Y, dtype=np.float32), device=dev).permute(2, 0, 1)[None].contiguous()
    Ct = torch.as_tensor(np.ascontiguousarray(crm, dtype=np.float32), device=dev).permute(2, 0, 1)[None].contiguous()
    return ops.icrm_forward(Yt, Ct)[0].permute(1, 2, 0).contiguous().cpu().numpy()
