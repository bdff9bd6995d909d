"""Objective waveform metrics of the evaluation step (TEST INFRASTRUCTURE): numpy restatement of
M2/metrics.py:86-129 (metrics_ssnr) and :132-175 (metrics_ssnr_shift) -- segmental SNR over 30 ms
Hann-windowed frames at a quarter-frame hop, each clamped to [min_snr, max_snr], plus the overall SNR.
Pinned by tests/golden/metrics.npz (oracle/make_golden_metrics.py runs the reference's own source)."""
import numpy as np


def _ssnr(ref_wav, deg_wav, srate, win_len, min_snr, max_snr, eps, shift):
    ref_wav = np.asarray(ref_wav, dtype=np.float64)
    deg_wav = np.asarray(deg_wav, dtype=np.float64)
    dif = ref_wav - deg_wav
    overall = 10 * np.log10(np.sum(ref_wav ** 2) / (np.sum(dif ** 2) + eps))           # metrics.py:97-98
    winlength = int(np.round(win_len * srate / 1000))                                   # :101
    skiprate = winlength // 4                                                           # :102
    num_frames = int(ref_wav.shape[0] / skiprate - (winlength / skiprate))              # :108
    time = np.linspace(1, winlength, winlength) / (winlength + 1)                       # :110
    window = 0.5 * (1 - np.cos(2 * np.pi * time))                                       # :111
    seg = []
    for f in range(num_frames):
        c = ref_wav[f * skiprate:f * skiprate + winlength] * window
        p = deg_wav[f * skiprate:f * skiprate + winlength] * window
        se, ne = np.sum(c ** 2), np.sum((c - p) ** 2)
        v = 10 * np.log10(se / (ne + eps) + (1.0 if shift else eps))                    # :125 / :171
        seg.append(min(max(v, min_snr), max_snr))
    return float(overall), float(np.nanmean(seg))


def metrics_ssnr(ref_wav, deg_wav, srate=16000, win_len=30, min_snr=-10, max_snr=35, eps=1e-10):
    return _ssnr(ref_wav, deg_wav, srate, win_len, min_snr, max_snr, eps, False)


def metrics_ssnr_shift(ref_wav, deg_wav, srate=16000, win_len=30, min_snr=-10, max_snr=35, eps=1e-10):
    return _ssnr(ref_wav, deg_wav, srate, win_len, min_snr, max_snr, eps, True)
