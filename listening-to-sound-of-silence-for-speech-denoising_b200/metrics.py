"""Objective waveform metrics of the evaluation step on the device, with the reference's names and argument
meaning (M2/metrics.py).  Built: the segmental / overall SNR family, the L1 distance, the weighted spectral slope (WSS) and
log-likelihood-ratio (LLR) measures and the composite scores with the PESQ value supplied by the caller (PESQ / STOI themselves are
third-party packages, pypesq / pystoi, absent here).

    metrics_ssnr(ref, deg, srate=16000, ...)        M2/metrics.py:86-129    -> (overall_snr, segmental_snr)
    metrics_ssnr_shift(...)                         M2/metrics.py:132-175
    metrics_L1(output, target)                      M2/metrics.py:40-45     (equal lengths: mean |output - target|)
    wss(ref, deg, srate, eps) / llr(ref, deg, srate) M2/metrics.py:404-558 / 561-681 -> per-frame distances (numpy array, or (B, frames))
    CompositeEval(ref, deg, srate, eps, pesq_raw=)  M2/metrics.py:346-401   -> Csig, Cbak, Covl, pesq_raw, segSNR, overall_snr

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


def _frame_metric(fn, ref_wav, deg_wav, srate, *extra):
    ref, single = _wave(ref_wav)
    deg, _ = _wave(deg_wav)
    if ref.shape != deg.shape:
        raise AssertionError(ref.shape[1])                          # the reference asserts equal lengths
    B, L = ref.shape
    nf = lib().sos_metric_frames(L, int(srate))
    if nf <= 0:
        return np.zeros(0) if single else torch.zeros(B, 0, dtype=torch.float64, device=ref.device)
    out = torch.empty(B, nf, device=ref.device, dtype=torch.float64)
    check(fn(ops._p(ref), ops._p(deg), B, L, int(srate), *extra, C.c_void_p(out.data_ptr()), ops._stream()), "sos_wss / sos_llr")
    ops._count()
    return out[0].cpu().numpy() if single else out


def wss(ref_wav, deg_wav, srate, eps=1e-10):
    return _frame_metric(lib().sos_wss, ref_wav, deg_wav, srate, float(eps))


def llr(ref_wav, deg_wav, srate):
    return _frame_metric(lib().sos_llr, ref_wav, deg_wav, srate)


def CompositeEval(ref_wav, deg_wav, srate=16000, eps=1e-10, pesq_raw=None):
    """M2/metrics.py:346-401 for one waveform pair.  The reference calls pypesq for `pesq_raw`; that third-party package is absent,
    so the caller supplies the value (None -> the three composite scores are None; segSNR / overall_snr are still returned)."""
    alpha = 0.95
    n = min(len(ref_wav), len(deg_wav))
    ref_wav, deg_wav = ref_wav[:n], deg_wav[:n]
    w = sorted(wss(ref_wav, deg_wav, srate, eps=eps))
    wss_dist = np.nanmean(w[:int(round(len(w) * alpha))])
    l = sorted(llr(ref_wav, deg_wav, srate))
    llr_mean = np.nanmean(l[:round(len(l) * alpha)])
    overall_snr, seg = metrics_ssnr(ref_wav, deg_wav, srate=srate, min_snr=0, eps=eps)
    if pesq_raw is None:
        return None, None, None, None, seg, overall_snr
    trim = lambda v: min(max(v, 1), 5)
    csig = trim(3.093 - 1.029 * llr_mean + 0.603 * pesq_raw - 0.009 * wss_dist)
    cbak = trim(1.634 + 0.478 * pesq_raw - 0.007 * wss_dist + 0.063 * seg)
    covl = trim(1.594 + 0.805 * pesq_raw - 0.512 * llr_mean - 0.007 * wss_dist)
    return csig, cbak, covl, pesq_raw, seg, overall_snr
