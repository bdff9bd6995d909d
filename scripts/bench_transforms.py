"""STFT / iSTFT kernels alone: time per launch (CUDA events, L2 flushed between launches by a 256 MB memset) and achieved fraction
of the measured HBM copy peak.  Algorithmic bytes: STFT 4 L + 4*2*256*T per signal, iSTFT 4*2*256*T + 4*158*(T-1).

    python scripts/bench_transforms.py [n_signals ...]        (SOS_STFT_TF32=1: the previous one-tile-per-CTA TF32 kernel)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sos_b200  # noqa: E402,F401
from sos_b200 import ops, transform  # noqa: E402


def timed(fn, flush, reps=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    ops.init()
    dev = torch.device("cuda:0")
    peak = 6551.0
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        pass
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    L = 32000
    T = 1 + L // 158
    for n in [int(a) for a in sys.argv[1:]] or [32, 128, 512]:
        wave = torch.randn(n, L, device=dev) * 0.2
        spec = transform.stft_batch(wave)
        ms = timed(lambda: ops.stft(wave), flush)
        by = 4.0 * n * L + 4.0 * n * 2 * 256 * T
        print(f"stft   {n:4d} signals: {1e3 * ms:8.1f} us  {by / ms / 1e6:7.1f} GB/s  {by / ms / 1e6 / peak:.3f} of the HBM copy peak ({peak:.0f} GB/s)")
        ms = timed(lambda: ops.istft(spec), flush)
        by = 4.0 * n * 2 * 256 * T + 4.0 * n * 158 * (T - 1)
        print(f"istft  {n:4d} signals: {1e3 * ms:8.1f} us  {by / ms / 1e6:7.1f} GB/s  {by / ms / 1e6 / peak:.3f} of the HBM copy peak")


if __name__ == "__main__":
    main()
