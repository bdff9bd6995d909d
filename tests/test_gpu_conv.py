"""GPU parity of the tcgen05 tap GEMM (forward / data gradient / weight gradient) against the oracle's convolutions.

The kernels multiply TF32 operands (10-bit mantissa, rounded to nearest where they are produced) and accumulate in fp32;
oracle.nets restates exactly that contract on the CPU (operands rounded with the same rule, fp32 convolution), so the
comparison is tight: TOL = 2e-5 of the output scale (accumulation order only).  Against plain fp32 convolutions the same
outputs differ by ~3e-4 of scale (TF32_TOL), which is also asserted."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-5
TF32_TOL = 4e-3


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _ref_conv(x, w, kind, k, d, stride):
    """x NCHW cpu, returns NCHW (in the oracle's current arithmetic contract)."""
    from oracle import nets
    if kind == "zero":
        pad = ((k[0] - 1) // 2 * d[0], (k[1] - 1) // 2 * d[1])
        return nets._conv(x, w, 1, pad, d)
    if kind == "valid":
        return nets._conv(x, w, stride, 0, d)
    return nets._convT(x, w)


CASES = [
    # kind, Cin, Cout, k, d, stride, N, H, W
    ("zero", 32, 32, (1, 1), (1, 1), 1, 1, 16, 24),
    ("zero", 48, 48, (5, 5), (1, 1), 1, 2, 32, 27),
    ("zero", 96, 96, (5, 5), (2, 2), 1, 1, 32, 43),
    ("zero", 96, 96, (5, 5), (1, 1), 1, 3, 16, 16),          # three tiles: the CTA-pair kernel's odd last tile gets a dummy partner
    ("zero", 96, 96, (5, 5), (8, 1), 1, 1, 64, 19),
    ("zero", 48, 48, (5, 5), (16, 16), 1, 1, 64, 37),
    ("zero", 2, 96, (1, 7), (1, 1), 1, 2, 16, 29),
    ("zero", 96, 96, (7, 1), (1, 1), 1, 1, 32, 21),
    ("zero", 96, 8, (1, 1), (1, 1), 1, 2, 16, 21),
    ("zero", 48, 4, (1, 1), (1, 1), 1, 1, 16, 21),
    ("valid", 2, 64, (5, 5), (1, 1), 1, 1, 36, 27),
    ("zero", 2, 48, (3, 3), (2, 1), 1, 2, 40, 33),           # two real input channels: the folded-tap form in half mode (16 / 32 / 64 columns)
    ("zero", 2, 96, (5, 5), (1, 2), 1, 1, 130, 37),
    ("valid", 64, 128, (5, 5), (1, 1), 2, 2, 36, 31),
    ("valid", 128, 128, (5, 5), (1, 1), 1, 1, 20, 23),
    ("valid", 256, 256, (3, 3), (1, 1), 2, 1, 34, 29),
    ("valid", 256, 256, (3, 3), (4, 4), 1, 1, 24, 21),
    ("valid", 64, 2, (3, 3), (1, 1), 1, 1, 18, 23),
    ("convT", 256, 128, (3, 3), (1, 1), 2, 1, 16, 13),
    ("convT", 128, 64, (3, 3), (1, 1), 2, 2, 8, 11),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[1]}to{c[2]}-k{c[3][0]}x{c[3][1]}-d{c[4][0]}x{c[4][1]}-s{c[5]}")
def test_tapconv_fwd_bwd(cuda, case):
    from sos_b200 import layers as L, ops
    kind, Cin, Cout, k, d, stride, N, H, W = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn(N, Cin, H, W, generator=g)
    if kind == "convT":
        w = torch.randn(Cin, Cout, 3, 3, generator=g) / (Cin * 9) ** 0.5
    else:
        w = torch.randn(Cout, Cin, k[0], k[1], generator=g) / (Cin * k[0] * k[1]) ** 0.5
    from oracle import nets
    x.requires_grad_(True)
    w.requires_grad_(True)
    y32 = _ref_conv(x, w, kind, k, d, stride).detach()                    # plain fp32
    with nets.tf32_contract():
        y_ref = _ref_conv(x, w, kind, k, d, stride)
    gy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(gy)

    geom = L.ConvGeom(kind, k[0], k[1], d[0], d[1], stride, round_dy=True)
    cin_p = (Cin + 7) // 8 * 8
    xd = ops.round_tf32_(ops.nchw_to_nhwc(x.detach().to(cuda), cin_p)).requires_grad_(True)
    wd = w.detach().to(cuda).requires_grad_(True)
    y = L.TapConv.apply(xd, wd, geom)
    torch.cuda.synchronize()
    y_nchw = ops.nhwc_to_nchw(y, Cout).cpu()
    assert y_nchw.shape == y_ref.shape
    e_f = _rel(y_nchw, y_ref.detach())
    gyd = ops.nchw_to_nhwc(gy.to(cuda), y.shape[3])
    y.backward(gyd)
    torch.cuda.synchronize()
    e_w = _rel(wd.grad.cpu(), w.grad)
    dx = ops.nhwc_to_nchw(xd.grad, Cin).cpu()
    e_x = _rel(dx, x.grad)
    print(f"\n{case}: fwd {e_f:.2e} dgrad {e_x:.2e} wgrad {e_w:.2e}")
    assert e_f < TOL and e_x < TOL and e_w < TOL, f"rel err: forward {e_f:.2e} dgrad {e_x:.2e} wgrad {e_w:.2e}"
    assert _rel(y_nchw, y32) < TF32_TOL


@pytest.mark.parametrize("plan", [0, 1, 2, 3])
@pytest.mark.parametrize("d", [(1, 1), (4, 4), (2, 1)])
def test_conv_forced_plans(cuda, plan, d):
    """Every orientation / box-sharing plan of the forward kernel must give the same answer."""
    from sos_b200 import layers as L, ops
    Cin, Cout, N, H, W = 48, 48, 1, 32, 40
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 5, 5, generator=g) / (Cin * 25) ** 0.5
    from oracle import nets
    with nets.tf32_contract():
        y_ref = nets._conv(x, w, 1, (2 * d[0], 2 * d[1]), d)
    geom = L.ConvGeom("zero", 5, 5, d[0], d[1], 1)
    xd = ops.round_tf32_(ops.nchw_to_nhwc(x.to(cuda), Cin))
    wk = L._pack_fwd(w.to(cuda), geom.taps, Cin)
    info = [0] * 8
    try:
        y = ops.conv_tc(xd, wk, [o[0] for o in geom.off], [o[1] for o in geom.off], Cout, H, W, 1, force_plan=plan, plan_out=info)
    except Exception as e:                              # a plan may be inapplicable (e.g. lattice does not divide the axis)
        if "not applicable" in str(e):
            pytest.skip(str(e))
        raise
    torch.cuda.synchronize()
    e = _rel(ops.nhwc_to_nchw(y, Cout).cpu(), y_ref)
    print(f"\nplan {plan} d {d}: info {info} err {e:.2e}")
    assert e < TOL


def test_conv_fused_epilogue(cuda):
    from sos_b200 import layers as L, ops
    Cin, Cout, N, H, W = 96, 96, 2, 32, 27
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 5, 5, generator=g) / (Cin * 25) ** 0.5
    sc, sh = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.1
    slope = torch.tensor([0.25])
    from oracle import nets
    with nets.tf32_contract():
        y_ref = nets._conv(x, w, 1, 2, 1) * sc[None, :, None, None] + sh[None, :, None, None]
    geom = L.ConvGeom("zero", 5, 5, 1, 1, 1)
    xd = ops.round_tf32_(ops.nchw_to_nhwc(x.to(cuda), Cin))
    for act, ref in ((1, torch.relu(y_ref)), (2, torch.where(y_ref > 0, y_ref, 0.25 * y_ref)), (0, y_ref)):
        y = L.conv_fused_eval(xd, w.to(cuda), geom, sc.to(cuda), sh.to(cuda), act, slope.to(cuda), round_out=False)
        torch.cuda.synchronize()
        e = _rel(ops.nhwc_to_nchw(y, Cout).cpu(), ref)
        assert e < TOL, (act, e)


# ------------------------------------------------------------------------------------------------ half operands (kind::f16)
# IEEE half has the same 11-bit significand as TF32, so (inside its exponent range) the half path obeys the same arithmetic
# contract up to the tie rule (cvt.rn.f16 rounds ties to even, cvt.rna.tf32 away from zero): oracle.nets.half_contract.  Gradient operands are far below
# half's normal range; they travel scaled by a per-tensor power of two (exact), which the tiny-gradient case exercises.
HALF_TOL = TOL             # vs the oracle's half contract (round-to-nearest-even operands, fp32 accumulate)
HALF_CASES = [c for c in CASES if c[1] % 16 == 0 or c[1] == 2]


@pytest.mark.parametrize("gscale", [1.0, 3e-9])
@pytest.mark.parametrize("case", HALF_CASES, ids=lambda c: f"{c[0]}-{c[1]}to{c[2]}-k{c[3][0]}x{c[3][1]}-d{c[4][0]}x{c[4][1]}-s{c[5]}")
def test_tapconv_half_fwd_bwd(cuda, case, gscale):
    from sos_b200 import layers as L, ops
    from oracle import nets
    kind, Cin, Cout, k, d, stride, N, H, W = case
    if gscale != 1.0 and (Cin, Cout) not in ((96, 96), (48, 48), (64, 128), (128, 64)):
        pytest.skip("tiny-gradient scaling is checked on a subset")
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn(N, Cin, H, W, generator=g)
    if kind == "convT":
        w = torch.randn(Cin, Cout, 3, 3, generator=g) / (Cin * 9) ** 0.5
    else:
        w = torch.randn(Cout, Cin, k[0], k[1], generator=g) / (Cin * k[0] * k[1]) ** 0.5
    x.requires_grad_(True)
    w.requires_grad_(True)
    with nets.half_contract():
        y_ref = _ref_conv(x, w, kind, k, d, stride)
        gy = torch.randn(y_ref.shape, generator=g) * gscale
        y_ref.backward(gy)

    geom = L.ConvGeom(kind, k[0], k[1], d[0], d[1], stride, round_dy=True)
    cin_p = (Cin + 15) // 16 * 16
    x32 = ops.nchw_to_nhwc(x.detach().to(cuda), cin_p).requires_grad_(True)
    xd = L.ToHalf.apply(x32, cin_p)
    assert ops.is_half_handle(xd) and ops.hv(xd).dtype == torch.float16
    wd = w.detach().to(cuda).requires_grad_(True)
    y = L.TapConvH.apply(xd, wd, geom)
    torch.cuda.synchronize()
    y_nchw = ops.nhwc_to_nchw(y, Cout).cpu()
    assert y_nchw.shape == y_ref.shape
    e_f = _rel(y_nchw, y_ref.detach())
    y.backward(ops.nchw_to_nhwc(gy.to(cuda), y.shape[3]))
    torch.cuda.synchronize()
    e_w = _rel(wd.grad.cpu(), w.grad)
    e_x = _rel(ops.nhwc_to_nchw(x32.grad, Cin).cpu(), x.grad)
    print(f"\nhalf {case} gscale {gscale}: fwd {e_f:.2e} dgrad {e_x:.2e} wgrad {e_w:.2e}")
    assert e_f < HALF_TOL and e_x < HALF_TOL and e_w < HALF_TOL, f"rel err: forward {e_f:.2e} dgrad {e_x:.2e} wgrad {e_w:.2e}"


@pytest.mark.parametrize("Cin,Cout,k", [(48, 64, (3, 3)), (96, 96, (5, 5)), (24, 32, (3, 3)), (48, 8, (1, 1))])
def test_conv_never_reads_past_its_weights(cuda, Cin, Cout, k):
    """The operand chunks of the last tap / last weight row reach past the packed weight matrix when Cin is not a multiple of the
    chunk width: whatever lies there must not enter the sums (0 x NaN = NaN).  The packed weights sit in the middle of a NaN-filled
    buffer; results must equal those of a private copy."""
    from sos_b200 import layers as L, ops
    g = torch.Generator().manual_seed(11)
    geom = L.ConvGeom("zero", k[0], k[1], 1, 1, 1)
    x = ops.to_half(torch.randn(2, 40, 36, Cin, generator=g).to(cuda))
    w = (torch.randn(Cout, Cin, k[0], k[1], generator=g) * 0.1).to(cuda)
    wk = ops.pack_taps_half(w, geom.taps, Cin)
    big = torch.full((wk.numel() + 16384,), float("nan"), device=cuda, dtype=torch.float16)
    wk2 = big[8192:8192 + wk.numel()].view_as(wk)
    wk2.copy_(wk)
    dh, dw = [o[0] for o in geom.off], [o[1] for o in geom.off]
    y1 = ops.conv_tc(x, wk, dh, dw, Cout, 40, 36, 1, y_half=True)
    y2 = ops.conv_tc(x, wk2, dh, dw, Cout, 40, 36, 1, y_half=True)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(y2.float()).all()) and torch.equal(y1, y2)


def test_im2col_half_kernel(cuda):
    """sos_im2col_half against an index-level restatement: column 2 t + c = x[n, oh + dh_t, ow + dw_t, c], zeros outside the image
    and in the padding columns; only the first two of the stored channels are read."""
    from sos_b200 import ops
    g = torch.Generator().manual_seed(4)
    for (N, H, W, Cp, offs, OH, OW, Kc) in [(2, 9, 13, 16, [(0, b - 3) for b in range(7)], 9, 13, 16),
                                            (1, 12, 10, 8, [(a, b) for a in range(5) for b in range(5)], 8, 6, 64),
                                            (3, 7, 5, 16, [(2 * (a - 1), b - 1) for a in range(3) for b in range(3)], 7, 5, 32),
                                            (2, 6, 150, 16, [(a - 2, b - 2) for a in range(5) for b in range(5)], 6, 150, 64),       # three column tiles
                                            (1, 40, 70, 8, [(4 * (a - 1), 16 * (b - 1)) for a in range(3) for b in range(3)], 40, 70, 32),
                                            (1, 12, 9, 16, [(a, 0) for a in range(9)], 4, 9, 32)]:       # 9 row offsets: the gather form
        x = torch.randn(N, H, W, Cp, generator=g).half()
        out = ops.im2col_half(x.to(cuda), [o[0] for o in offs], [o[1] for o in offs], OH, OW, Kc).cpu()
        ref = torch.zeros(N, OH, OW, Kc, dtype=torch.float16)
        for t, (dh, dw) in enumerate(offs):
            for oh in range(OH):
                for ow in range(OW):
                    if 0 <= oh + dh < H and 0 <= ow + dw < W:
                        ref[:, oh, ow, 2 * t:2 * t + 2] = x[:, oh + dh, ow + dw, :2]
        assert torch.equal(out, ref)


def test_folded_taps_match_tap_form(cuda):
    """The folded form and the tap form of a two-channel convolution multiply the same operand values: forward and weight gradient
    agree to accumulation order."""
    from sos_b200 import layers as L, ops
    g = torch.Generator().manual_seed(6)
    x = ops.to_half(torch.randn(2, 64, 45, 16, generator=g).to(cuda))
    w = (torch.randn(64, 2, 5, 5, generator=g) * 0.1).to(cuda)
    geom = L.ConvGeom("zero", 5, 5, 1, 1, 1)
    dy = ops.to_half(torch.randn(2, 64, 45, 64, generator=g).to(cuda))
    res = {}
    for fold in (True, False):
        old, L._FOLD_TAPS, old_f, L._FOLD_FWD = L._FOLD_TAPS, fold, L._FOLD_FWD, fold
        try:
            assert bool(L._fold_kc(x, w, geom)) == fold
            res[fold] = (L._conv_forward(x, w, geom).clone(), L._conv_wgrad(x, dy, w, geom).clone())
        finally:
            L._FOLD_TAPS, L._FOLD_FWD = old, old_f
    assert _rel(res[True][0], res[False][0]) < 1e-5 and _rel(res[True][1], res[False][1]) < 1e-5


def test_conv_half_fused_epilogue(cuda):
    """Inference epilogue with half outputs written straight into the next layer's operand map."""
    from sos_b200 import layers as L, ops
    from oracle import nets
    g = torch.Generator().manual_seed(11)
    for Cin, Cout, kind, k in ((96, 96, "zero", 5), (48, 48, "zero", 5), (64, 128, "valid", 5), (256, 128, "convT", 3)):
        N, H, W = 2, 32, 27
        x = torch.randn(N, Cin, H, W, generator=g)
        w = (torch.randn(Cin, Cout, 3, 3, generator=g) if kind == "convT" else torch.randn(Cout, Cin, k, k, generator=g)) / (Cin * k * k) ** 0.5
        sc, sh = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.1
        with nets.half_contract():
            y_ref = _ref_conv(x, w, kind, (k, k), (1, 1), 1) * sc[None, :, None, None] + sh[None, :, None, None]
            ref = nets.round_tf32(torch.relu(y_ref))
        geom = L.ConvGeom(kind, k, k, 1, 1, 2 if kind == "convT" else 1)
        xd = L.ToHalf.apply(ops.nchw_to_nhwc(x.to(cuda), Cin), Cin)
        z = L.conv_fused_eval(xd, w.to(cuda), geom, sc.to(cuda), sh.to(cuda), 1, None, round_out=True)
        torch.cuda.synchronize()
        assert ops.is_half_handle(z)
        got = ops.hv(z).float().permute(0, 3, 1, 2).cpu()
        e = _rel(got, ref)
        print(f"\nhalf epilogue {kind} {Cin}->{Cout}: {e:.2e}")
        assert e < 1.1e-3, (kind, Cin, Cout, e)        # one ulp of the half output (<= 2^-10 of its value) on top of the GEMM tolerance
