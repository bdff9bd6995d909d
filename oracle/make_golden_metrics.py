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


if __name__ == "__main__":
    ssnr, ssnr_shift = reference_functions()
    rows = []
    for index, length, clean, mixed in cases():
        for eps in (1e-10, 1e-20):
            a = ssnr(clean.astype(np.float64), mixed.astype(np.float64), eps=eps)
            b = ssnr_shift(clean.astype(np.float64), mixed.astype(np.float64), eps=eps)
            rows.append([index, length, eps, a[0], a[1], b[0], b[1]])
    np.savez(os.path.join(ROOT, "tests", "golden", "metrics.npz"), rows=np.array(rows, dtype=np.float64))
    print(np.array(rows))
