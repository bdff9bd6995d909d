"""GPU parity of the row-streaming stacked-tap kernel (csrc/conv_row.cu: the 48-channel dilated 5x5 layers, forward with the BatchNorm
statistics in the epilogue and data gradient, half outputs) against the oracle's convolution under the half-operand contract.
plan_out[0] == 2 proves that the row kernel (not the tap-list GEMM) served the call."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


CASES = [  # dilation (dh, dw), N, H, W
    ((1, 1), 2, 256, 37), ((2, 1), 1, 128, 45), ((4, 1), 1, 256, 23), ((2, 2), 2, 128, 41), ((4, 4), 1, 256, 53),
    ((8, 8), 1, 128, 61), ((16, 1), 1, 256, 19), ((16, 16), 1, 256, 67), ((1, 1), 1, 128, 203), ((8, 1), 3, 128, 9),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"d{c[0][0]}x{c[0][1]}-n{c[1]}-{c[2]}x{c[3]}")
def test_rowconv_forward_stats_and_dgrad(cuda, case):
    from oracle import nets
    from sos_b200 import layers as L, ops
    d, N, H, W = case
    C = 48
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(C, C, 5, 5, generator=g) / (C * 25) ** 0.5
    gy = torch.randn(N, C, H, W, generator=g)
    x.requires_grad_(True)
    with nets.half_contract():
        y_ref = nets._conv(x, w, 1, (2 * d[0], 2 * d[1]), d)
        y_ref.backward(gy)
    geom = L.ConvGeom("zero", 5, 5, d[0], d[1], 1)
    xh = ops.to_half(ops.nchw_to_nhwc(x.detach().to(cuda), C))
    wd = w.to(cuda)
    # ---- forward: half output + per-channel sums of the fp32 accumulators
    wk = L._pack_fwd(wd, geom.taps, C, True)
    info = [0] * 8
    y, partial = ops.conv_tc(xh, wk, [o[0] for o in geom.off], [o[1] for o in geom.off], C, H, W, 1, want_stats=True, y_half=True, plan_out=info)
    torch.cuda.synchronize()
    assert info[0] == 2, f"the row-streaming kernel did not serve this call: {info}"
    got = y.float().permute(0, 3, 1, 2).cpu()
    e_f = _rel(got, y_ref.detach())
    sums = partial.sum(0).cpu()                                                    # (2, C)
    ref_s, ref_q = y_ref.detach().sum((0, 2, 3)), (y_ref.detach() ** 2).sum((0, 2, 3))
    e_s = float((sums[0] - ref_s).abs().max() / (ref_q.sqrt().max() + 1e-12))
    e_q = _rel(sums[1], ref_q)
    # ---- data gradient through the same kernel (flipped taps, transposed weights), stored as half
    gyh = ops.to_half(ops.nchw_to_nhwc(gy.to(cuda), C))
    info2 = [0] * 8
    wt = wd.permute(1, 0, 2, 3)
    wk2 = L._pack_fwd(wt, geom.taps, C, True)
    dx = ops.conv_tc(gyh, wk2, [-o[0] for o in geom.off], [-o[1] for o in geom.off], C, H, W, 1, y_half=True, plan_out=info2)
    torch.cuda.synchronize()
    assert info2[0] == 2
    e_x = _rel(dx.float().permute(0, 3, 1, 2).cpu(), x.grad)
    print(f"\nrowconv {case}: fwd {e_f:.2e} sum {e_s:.2e} sumsq {e_q:.2e} dgrad {e_x:.2e}")
    # half outputs: one rounding (2^-11 of the value) on top of the GEMM's accumulation-order tolerance
    assert e_f < 6e-4 and e_x < 6e-4, (e_f, e_x)
    assert e_s < 2e-5 and e_q < 2e-5, (e_s, e_q)


def test_rowconv_eval_epilogue(cuda):
    """Inference form: BatchNorm affine + ReLU in the epilogue, half output feeding the next layer."""
    from oracle import nets
    from sos_b200 import layers as L, ops
    g = torch.Generator().manual_seed(3)
    N, C, H, W = 2, 48, 128, 33
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(C, C, 5, 5, generator=g) / (C * 25) ** 0.5
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    with nets.half_contract():
        ref = torch.relu(nets._conv(x, w, 1, (4, 4), (2, 2)) * sc[None, :, None, None] + sh[None, :, None, None])
    geom = L.ConvGeom("zero", 5, 5, 2, 2, 1)
    xd = L.ToHalf.apply(ops.nchw_to_nhwc(x.to(cuda), C), C)
    z = L.conv_fused_eval(xd, w.to(cuda), geom, sc.to(cuda), sh.to(cuda), 1, None, round_out=True)
    torch.cuda.synchronize()
    e = _rel(ops.hv(z).float().permute(0, 3, 1, 2).cpu(), ref)
    print(f"\nrowconv eval epilogue: {e:.2e}")
    assert e < 1.1e-3


@pytest.mark.parametrize("d", [(1, 1), (4, 4), (8, 1)])
def test_rowconv_fused_bn_backward_reduction(cuda, d):
    """A data-gradient call inside an encoder chain also delivers pass 1 of the BatchNorm backward of the layer below (sos_conv_args::
    bnr_*): per-channel sums of g, g * xhat and g^2 over the dz it writes, g = dz where that layer's ReLU was active.  Checked against
    the same sums taken from the stored half dz, and end to end: sos_bn_act_backward_half_pre == sos_bn_act_backward_half."""
    from sos_b200 import layers as L, ops
    N, C, H, W = 2, 48, 128, 39
    g = torch.Generator().manual_seed(7 + d[0])
    dy = ops.to_half((torch.randn(N, H, W, C, generator=g) * 0.5).to(cuda))
    w = (torch.randn(C, C, 5, 5, generator=g) / (C * 25) ** 0.5).to(cuda)
    y_below = ops.to_half((torch.randn(N, H, W, C, generator=g) * 1.5 + 0.3).to(cuda))
    mean, var = torch.randn(C, generator=g).to(cuda) * 0.2, (torch.rand(C, generator=g) + 0.5).to(cuda)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(cuda), (torch.randn(C, generator=g) * 0.3).to(cuda)
    invstd = torch.rsqrt(var + 1e-5)
    scale = gamma * invstd
    shift = beta - mean * scale
    stats = [mean.contiguous(), invstd.contiguous(), scale.contiguous(), shift.contiguous()]
    geom = L.ConvGeom("zero", 5, 5, d[0], d[1], 1)
    dz, part = L._conv_dgrad_raw(dy, w, geom, (N, H, W, C), None, y_half=True, bnr=(y_below, stats))
    torch.cuda.synchronize()
    assert part is not None and part.shape[1:] == (4, C), "the row-streaming kernel did not deliver the fused reduction"
    dzf, yf = dz.float().reshape(-1, C), y_below.float().reshape(-1, C)
    gm = torch.where(yf * scale + shift > 0, dzf, torch.zeros_like(dzf)).double()
    xhat = ((yf - mean) * invstd).double()
    want = torch.stack([gm.sum(0), (gm * xhat).sum(0), torch.zeros(C, device=cuda, dtype=torch.float64), (gm * gm).sum(0)])
    got = part.double().sum(0)
    err = float((got - want).abs().max() / want.abs().max())
    print(f"\nfused bn-backward reduction d={d}: rel err {err:.2e} over {part.shape[0]} partial rows")
    assert err < 2e-5
    # end to end: backward of the layer below from the fused partial sums == from its own reduction pass
    inv = torch.ones(1, device=cuda)
    a = ops.bn_train_backward_half(dz, y_below, stats, ops.ACT_RELU, None, dz_inv=inv, pre_partial=part)
    b = ops.bn_train_backward_half(dz, y_below, stats, ops.ACT_RELU, None, dz_inv=inv)
    torch.cuda.synchronize()
    assert float((a[0].float() - b[0].float()).abs().max()) <= 2e-3 * float(b[0].float().abs().max())
    for i in (1, 2):
        assert float((a[i] - b[i]).abs().max()) <= 1e-4 * float(b[i].abs().max()), i
    assert float((a[4][:2] - b[4][:2]).abs().max()) == 0.0          # the same power-of-two scale
