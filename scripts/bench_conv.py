"""Per-layer timing of the tcgen05 tap GEMM in its three roles (forward / data gradient / weight gradient) on the distinct
convolution shapes of the two networks at the benchmark size (B clips, T = 203).  Prints TFLOP/s (algorithmic 2*MACs) and
the planner's choice per layer.  usage: python scripts/bench_conv.py [B] [filter-substring] [tf32]   (default: half operands)"""
import sys
import torch
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import layers as L, ops

ops.init()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flt = sys.argv[2] if len(sys.argv) > 2 else ""
HALF = not (len(sys.argv) > 3 and sys.argv[3] == "tf32")
if flt == "all":
    flt = ""
T = 203
# name, kind, Cin, Cout, k, d, stride, H, W (input, already padded for "valid"), count in the step
LAYERS = [
    ("x96_k5_d1", "zero", 96, 96, (5, 5), (1, 1), 1, 256, T, 2), ("x96_k5_d2x1", "zero", 96, 96, (5, 5), (2, 1), 1, 256, T, 1),
    ("x96_k5_d8x1", "zero", 96, 96, (5, 5), (8, 1), 1, 256, T, 1), ("x96_k5_d32x1", "zero", 96, 96, (5, 5), (32, 1), 1, 256, T, 1),
    ("x96_k5_d2x2", "zero", 96, 96, (5, 5), (2, 2), 1, 256, T, 1), ("x96_k5_d8x8", "zero", 96, 96, (5, 5), (8, 8), 1, 256, T, 1),
    ("x96_k5_d32x32", "zero", 96, 96, (5, 5), (32, 32), 1, 256, T, 1), ("x96_k7x1", "zero", 96, 96, (7, 1), (1, 1), 1, 256, T, 1),
    ("x2to96_k1x7", "zero", 8, 96, (1, 7), (1, 1), 1, 256, T, 1), ("x96to8_k1", "zero", 96, 8, (1, 1), (1, 1), 1, 256, T, 1),
    ("n48_k5_d1", "zero", 48, 48, (5, 5), (1, 1), 1, 256, T, 4), ("n48_k5_d4x4", "zero", 48, 48, (5, 5), (4, 4), 1, 256, T, 2),
    ("n48_k5_d16x1", "zero", 48, 48, (5, 5), (16, 1), 1, 256, T, 2), ("n48_k5_d32x32", "zero", 48, 48, (5, 5), (32, 32), 1, 256, T, 1),
    ("in_2to64_k5", "valid", 8, 64, (5, 5), (1, 1), 1, 260, T + 4, 2), ("in_64to128_k5s2", "valid", 64, 128, (5, 5), (1, 1), 2, 260, T + 4, 2),
    ("in_128_k5", "valid", 128, 128, (5, 5), (1, 1), 1, 132, 106, 2), ("in_256_k3s2", "valid", 256, 256, (3, 3), (1, 1), 2, 130, 104, 1),
    ("in_256_k3", "valid", 256, 256, (3, 3), (1, 1), 1, 66, 53, 3), ("in_256_k3d4", "valid", 256, 256, (3, 3), (4, 4), 1, 72, 59, 1),
    ("in_256_k3d16", "valid", 256, 256, (3, 3), (16, 16), 1, 96, 83, 1), ("in_T256to128", "convT", 256, 128, (3, 3), (1, 1), 2, 64, 51, 1),
    ("in_256to128_k3", "valid", 256, 128, (3, 3), (1, 1), 1, 130, 104, 1), ("in_T128to64", "convT", 128, 64, (3, 3), (1, 1), 2, 128, 102, 1),
    ("in_128to64_k3", "valid", 128, 64, (3, 3), (1, 1), 1, 258, T + 2, 1), ("in_64to2_k3", "valid", 64, 2, (3, 3), (1, 1), 1, 258, T + 2, 1),
    # probes (count 0: not part of the step): the same small layers with their few channels stored as 16 / 32
    ("probe_64to16_k3", "valid", 64, 16, (3, 3), (1, 1), 1, 258, T + 2, 0), ("probe_64to32_k3", "valid", 64, 32, (3, 3), (1, 1), 1, 258, T + 2, 0),
    ("probe_96to32_k1", "zero", 96, 32, (1, 1), (1, 1), 1, 256, T, 0), ("probe_32to96_k1x7", "zero", 32, 96, (1, 7), (1, 1), 1, 256, T, 0),
]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
totf = 0.0
print(f"B={B}  {'layer':16s} {'GF':>8s} | {'fwd ms':>8s} {'TF/s':>6s} | {'dgrad ms':>8s} {'TF/s':>6s} | {'wgrad ms':>8s} {'TF/s':>6s}")
for name, kind, Cin, Cout, k, d, stride, H, W, cnt in LAYERS:
    if flt and flt not in name:
        continue
    g = L.ConvGeom(kind, k[0], k[1], d[0], d[1], stride)
    Cw = Cin
    if HALF and Cin == 8:
        Cin, Cw = 16, 2                  # the spectrogram inputs: 2 real channels (the weight's) in 16 stored ones -> folded taps (layers._fold_kc)
    x = torch.randn(B, H, W, Cin, device=dev).relu_()
    x = ops.to_half(x) if HALF else ops.round_tf32_(x)
    w = torch.randn((Cin, Cout, 3, 3) if kind == "convT" else (Cout, Cw, k[0], k[1]), device=dev) * 0.05
    OH, OW = g.out_size(H, W)
    y = L._conv_forward(x, w, g)
    dy = ops.to_half(torch.randn_like(y)) if HALF else ops.round_tf32_(torch.randn_like(y))
    flops = 2.0 * B * OH * OW * Cout * (2 if Cin == 16 and Cout in (96, 64) else Cin) * k[0] * k[1] / (4 if kind == "convT" else 1)
    if HALF:
        # the calls of a training step: forward with the BatchNorm statistics in the epilogue and a half output; data gradient
        # stored as half inside the encoder chains ("zero" layers), fp32 with the operand's inverse scale elsewhere
        inv = torch.ones(1, device=dev)
        t_f = timeit(lambda: L._conv_forward(x, w, g, want_stats=True, y_half=True))
        if kind == "zero":
            t_d = timeit(lambda: L._conv_dgrad_raw(dy, w, g, x.shape, None, y_half=True))
        else:
            t_d = timeit(lambda: L._conv_dgrad(dy, w, g, x.shape, inv))
    else:
        t_f = timeit(lambda: L._conv_forward(x, w, g))
        t_d = timeit(lambda: L._conv_dgrad(dy, w, g, x.shape))
    t_w = timeit(lambda: L._conv_wgrad_raw(x, dy, w, g, inv if HALF else None))
    tot["fwd"] += cnt * t_f; tot["dgrad"] += cnt * t_d; tot["wgrad"] += cnt * t_w; totf += cnt * flops
    print(f"      {name:16s} {flops/1e9:8.1f} | {t_f:8.3f} {flops/t_f/1e9:6.0f} | {t_d:8.3f} {flops/t_d/1e9:6.0f} | {t_w:8.3f} {flops/t_w/1e9:6.0f}  x{cnt}", flush=True)
    del x, w, y, dy
print(f"weighted totals (ms/step): fwd {tot['fwd']:.1f} dgrad {tot['dgrad']:.1f} wgrad {tot['wgrad']:.1f}; {totf/1e12:.2f} TFLOP per role")
