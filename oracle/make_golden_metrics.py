"""Golden values for the segmental-SNR metrics, produced by the REFERENCE's own source: the two pure-numpy functions are cut out
of M2/metrics.py (the module itself imports pypesq / pystoi / soundfile, which are absent here) and exec'd unmodified.
Run in the build container:  python -m oracle.make_golden_metrics   ->  tests/golden/metrics.npz"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/model_2_audio_denoising/audio_denoising_model/metrics.py"


def reference_functions():
    src = open(REF).read()
    ns = {"np": np}
    for name in ("metrics_ssnr", "metrics_ssnr_shift"):
        m = re.search(r"^def %s\(.*?(?=^def )" % name, src, flags=re.S | re.M)
        exec(m.group(0), ns)
    return ns["metrics_ssnr"], ns["metrics_ssnr_shift"]


def cases():
    sys.path.insert(0, ROOT)
    from oracle import synth
    out = []
    for index, length in ((0, 32000), (1, 28000), (2, 32000), (5, 16000)):
        c = synth.make_clip(index, length)
        out.append((index, length, c["clean"], c["mixed"]))
    return out


def reference_lpc_functions():
    """wss / llr / lpcoeff of M2/metrics.py, exec'd unmodified (they need numpy and scipy.linalg.toeplitz only)."""
    from scipy.linalg import toeplitz
    src = open(REF).read()
    ns = {"np": np, "toeplitz": toeplitz}
    for name in ("wss", "llr", "lpcoeff"):
        m = re.search(r"^def %s\(.*?(?=^(def |#@nb))" % name, src, flags=re.S | re.M) or re.search(r"^def %s\(.*" % name, src, flags=re.S | re.M)
        exec(m.group(0), ns)
    return ns["wss"], ns["llr"]


def lpc_cases():
    sys.path.insert(0, ROOT)
    from oracle import synth
    out = []
    for index, length, srate in ((0, 16000, 16000), (3, 14000, 14000), (6, 8000, 8000)):
        out.append((index, length, srate) + lpc_pair(index, length, srate))
    return out


def lpc_pair(index, length, srate):
    """Reference / degraded waveforms of the WSS / LLR fixtures: the synthetic clean speech (a sum of five harmonics, which an order-16
    predictor would fit almost perfectly and leave float32 cancellation noise in the LLR's quadratic forms) plus white noise at
    -26 dB, against the same speech plus a stronger, coloured noise."""
    sys.path.insert(0, ROOT)
    from oracle import synth
    c = synth.make_clip(index, length, srate)
    rng = np.random.default_rng(9000 + index)
    ref = c["clean"].astype(np.float64) + 0.01 * rng.standard_normal(length)
    deg = 0.9 * c["clean"].astype(np.float64) + 0.15 * c["full_noise"] + 0.02 * rng.standard_normal(length)
    return ref.astype(np.float32).astype(np.float64), deg.astype(np.float32).astype(np.float64)


if __name__ == "__main__":
    wss_ref, llr_ref = reference_lpc_functions()
    lpc = {}
    for index, length, srate, ref_w, deg_w in lpc_cases():
        lpc[f"wss:{index}:{length}:{srate}"] = np.array(wss_ref(ref_w, deg_w, srate), dtype=np.float64)
        lpc[f"llr:{index}:{length}:{srate}"] = np.array(llr_ref(ref_w, deg_w, srate), dtype=np.float64)
        print(index, srate, "wss frames", len(lpc[f"wss:{index}:{length}:{srate}"]), "mean", lpc[f"wss:{index}:{length}:{srate}"].mean(),
              "llr mean", np.nanmean(lpc[f"llr:{index}:{length}:{srate}"]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics_lpc.npz"), **lpc)
    ssnr, ssnr_shift = reference_functions()
    rows = []
    for index, length, clean, mixed in cases():
        for eps in (1e-10, 1e-20):
            a = ssnr(clean.astype(np.float64), mixed.astype(np.float64), eps=eps)
            b = ssnr_shift(clean.astype(np.float64), mixed.astype(np.float64), eps=eps)
            rows.append([index, length, eps, a[0], a[1], b[0], b[1]])
    np.savez(os.path.join(ROOT, "tests", "golden", "metrics.npz"), rows=np.array(rows, dtype=np.float64))
    print(np.array(rows))
