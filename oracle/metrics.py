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


# ---------------------------------------------------------------------------------------------- WSS / LLR / composite
# numpy restatement of M2/metrics.py:404-558 (wss), :561-623 (llr), :626-681 (lpcoeff) and the PESQ-free part of CompositeEval
# (:346-401).  Pinned by tests/golden/metrics_lpc.npz (oracle/make_golden_metrics.py runs the reference's own source).
_CENT = [50., 120, 190, 260, 330, 400, 470, 540, 617.372, 703.378, 798.717, 904.128, 1020.38, 1148.30, 1288.72, 1442.54, 1610.70,
         1794.16, 1993.93, 2211.08, 2446.71, 2701.97, 2978.04, 3276.17, 3597.63]                           # metrics.py:425-429
_BW = [70., 70, 70, 70, 70, 70, 70, 77.3724, 86.0056, 95.3398, 105.411, 116.256, 127.914, 140.423, 153.823, 168.154, 183.457,
       199.776, 217.153, 235.631, 255.255, 276.072, 298.126, 321.465, 346.136]                             # :430-434


def _frames(x, srate):
    winlength = int(round(30 * srate / 1000.))                                          # :412 / :569
    skiprate = int(np.floor(winlength / 4))
    num_frames = int(x.shape[0] / skiprate - (winlength / skiprate))                    # :459 / :579
    time = np.linspace(1, winlength, winlength) / (winlength + 1)
    window = 0.5 * (1 - np.cos(2 * np.pi * time))
    idx = np.arange(num_frames)[:, None] * skiprate + np.arange(winlength)[None, :]
    return x[idx] * window, winlength


def wss(ref_wav, deg_wav, srate, eps=1e-10):
    """-> per-frame weighted spectral slope distances (list semantics of the reference, as an array)."""
    ref_wav, deg_wav = np.asarray(ref_wav, np.float64), np.asarray(deg_wav, np.float64)
    cf, winlength = _frames(ref_wav, srate)
    pf, _ = _frames(deg_wav, srate)
    num_crit, max_freq = 25, srate / 2
    n_fft = int(2 ** np.ceil(np.log(2 * winlength) / np.log(2)))
    nby2 = n_fft // 2
    j = np.arange(nby2)
    filt = np.zeros((num_crit, nby2))
    min_factor = np.exp(-30. / (2 * 2.303))
    for i in range(num_crit):
        f0 = np.floor((_CENT[i] / max_freq) * nby2)
        bw = (_BW[i] / max_freq) * nby2
        f = np.exp(-11 * (((j - f0) / bw) ** 2) + (np.log(_BW[0]) - np.log(_BW[i])))
        filt[i] = f * (f > min_factor)
    out = []
    for c, p in zip(cf, pf):
        ce = 10 * np.log10(np.maximum(filt @ (np.abs(np.fft.fft(c, n_fft)) ** 2)[:nby2], eps))
        pe = 10 * np.log10(np.maximum(filt @ (np.abs(np.fft.fft(p, n_fft)) ** 2)[:nby2], eps))
        cs, ps = ce[1:] - ce[:-1], pe[1:] - pe[:-1]

        def peaks(en, sl):
            pk = []
            for i in range(num_crit - 1):
                n = i
                if sl[i] > 0:
                    while n < num_crit - 1 and sl[n] > 0:
                        n += 1
                    pk.append(en[n - 1])
                else:
                    while n >= 0 and sl[n] <= 0:
                        n -= 1
                    pk.append(en[n + 1])
            return np.array(pk)
        cpk, ppk = peaks(ce, cs), peaks(pe, ps)
        wc = (20 / (20 + ce.max() - ce[:-1])) * (1 / (1 + cpk - ce[:-1]))
        wp = (20 / (20 + pe.max() - pe[:-1])) * (1 / (1 + ppk - pe[:-1]))
        w = (wc + wp) / 2
        out.append(np.sum(w * (cs - ps) ** 2) / np.sum(w))
    return np.array(out)


def lpcoeff(frame, order):
    n = frame.shape[0]
    R = np.array([np.sum(frame[:n - k] * frame[k:]) for k in range(order + 1)])
    a = np.ones(order)
    E = np.zeros(order + 1)
    E[0] = R[0]
    for i in range(order):
        past = a[:i].copy()
        sum_term = np.sum(past * R[i:0:-1]) if i else 0.0
        rc = (R[i + 1] - sum_term) / E[i]
        a[i] = rc
        if i:
            a[:i] = past - rc * past[::-1]
        E[i + 1] = (1 - rc * rc) * E[i]
    return R.astype(np.float32), np.array([1.0] + list(-a), dtype=np.float32)


def llr(ref_wav, deg_wav, srate):
    from scipy.linalg import toeplitz
    ref_wav, deg_wav = np.asarray(ref_wav, np.float64), np.asarray(deg_wav, np.float64)
    cf, _ = _frames(ref_wav, srate)
    pf, _ = _frames(deg_wav, srate)
    P = 10 if srate < 10000 else 16
    out = []
    for c, p in zip(cf, pf):
        Rc, Ac = lpcoeff(c, P)
        _, Ap = lpcoeff(p, P)
        Ac, Ap = Ac[None, :], Ap[None, :]
        num = Ap.dot(toeplitz(Rc)).dot(Ap.T)
        den = Ac.dot(toeplitz(Rc)).dot(Ac.T)
        out.append(np.squeeze(np.log(num / den)))
    return np.array(out)


def composite_without_pesq(ref_wav, deg_wav, pesq_raw, srate=16000, eps=1e-10, alpha=0.95):
    """CompositeEval (M2/metrics.py:346-401) with the PESQ score supplied by the caller (pypesq is an absent third-party
    dependency): -> Csig, Cbak, Covl, segSNR, overall_snr, wss_dist, llr_mean."""
    n = min(len(ref_wav), len(deg_wav))
    ref_wav, deg_wav = np.asarray(ref_wav[:n], np.float64), np.asarray(deg_wav[:n], np.float64)
    w = sorted(wss(ref_wav, deg_wav, srate, eps=eps))
    wss_dist = np.nanmean(w[:int(round(len(w) * alpha))])
    l = sorted(llr(ref_wav, deg_wav, srate))
    llr_mean = np.nanmean(l[:round(len(l) * alpha)])
    overall_snr, seg = metrics_ssnr(ref_wav, deg_wav, srate=srate, min_snr=0, eps=eps)
    trim = lambda v: min(max(v, 1), 5)
    csig = trim(3.093 - 1.029 * llr_mean + 0.603 * pesq_raw - 0.009 * wss_dist)
    cbak = trim(1.634 + 0.478 * pesq_raw - 0.007 * wss_dist + 0.063 * seg)
    covl = trim(1.594 + 0.805 * pesq_raw - 0.512 * llr_mean - 0.007 * wss_dist)
    return csig, cbak, covl, seg, overall_snr, wss_dist, llr_mean
