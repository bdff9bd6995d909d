"""Objective waveform metrics of the evaluation step on the device, with the reference's names and argument
meaning (M2/metrics.py).  Built: the segmental / overall SNR family and the L1 distance; PESQ / STOI / the composite measures need
third-party code that is not part of the hot path.

    metrics_ssnr(ref, deg, srate=16000, ...)        M2/metrics.py:86-129    -> (overall_snr, segmental_snr)
    metrics_ssnr_shift(...)                         M2/metrics.py:132-175
    metrics_L1(output, target)                      M2/metrics.py:40-45     (equal lengths: mean |output - target|)

Waveforms may be 1-D (floats returned, like the reference) or (B, L) batches (tensors of B values returned)."""
import ctypes as C

import numpy as np
import torch

from . import ops
from ._lib import check, lib


def _wave(x):
    ops.init()
    t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)) if not torch.is_tensor(x) else x
    t = t.to(torch.device("cuda", torch.cuda.current_device()), dtype=torch.float32)
    return (t[None] if t.dim() == 1 else t).contiguous(), t.dim() == 1


def _ssnr(ref_wav, deg_wav, srate, win_len, min_snr, max_snr, eps, shift):
    ref, single = _wave(ref_wav)
    deg, _ = _wave(deg_wav)
    if ref.shape != deg.shape:
        raise ValueError(f"reference {tuple(ref.shape)} and degraded {tuple(deg.shape)} waveforms must have the same length")
    B, L = ref.shape
    out = torch.empty(2, B, device=ref.device, dtype=torch.float32)
    check(lib().sos_ssnr(ops._p(ref), ops._p(deg), B, L, int(srate), float(win_len), float(min_snr), float(max_snr), float(eps), int(shift),
                         C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), ops._stream()), "sos_ssnr")
    ops._count()
    if single:
        o = out.cpu()
        return float(o[0, 0]), float(o[1, 0])
    return out[0], out[1]


def metrics_ssnr(ref_wav, deg_wav, srate=16000, win_len=30, min_snr=-10, max_snr=35, eps=1e-10):
    return _ssnr(ref_wav, deg_wav, srate, win_len, min_snr, max_snr, eps, False)


def metrics_ssnr_shift(ref_wav, deg_wav, srate=16000, win_len=30, min_snr=-10, max_snr=35, eps=1e-10):
    return _ssnr(ref_wav, deg_wav, srate, win_len, min_snr, max_snr, eps, True)


def metrics_L1(output, target):
    out, single = _wave(output)
    tgt, _ = _wave(target)
    if out.shape != tgt.shape:
        raise NotImplementedError("metrics_L1 with resampling (different lengths) is not built")
    v = (out - tgt).abs().mean(dim=1)
    return float(v[0]) if single else v
