"""CPU emulation of the two tensor-core entry points' SEMANTICS (sos_conv2d_tc / sos_conv2d_wgrad as specified in
include/sos_b200.h), used only to test the host-side geometry (tap lists, sub-pixel phases, weight packing) without a GPU."""
import torch


def _gather(x, dh, dw, OH, OW, stride):
    N, H, W, C = x.shape
    oh = torch.arange(OH) * stride + dh
    ow = torch.arange(OW) * stride + dw
    vh, vw = (oh >= 0) & (oh < H), (ow >= 0) & (ow < W)
    g = x[:, oh.clamp(0, H - 1)][:, :, ow.clamp(0, W - 1)]
    return g * (vh[:, None] & vw[None, :])[None, :, :, None]


def conv_tc(x, wk, tap_dh, tap_dw, Cout, OH, OW, stride=1, y=None, y_coff=0, lattice=(1, 1, 0, 0), epi_scale=None, epi_shift=None, bnr=None,
            act=0, slope=None, force_plan=-1, plan_out=None, k_real=None, tag=None, want_stats=False, y_half=False, out_scale=None):
    N, H, W, Cin = x.shape
    osh, osw, oph, opw = lattice
    if y is None:
        y = torch.zeros(N, OH * osh, OW * osw, (Cout + 7) // 8 * 8)
    acc = torch.zeros(N, OH, OW, Cout)
    for t, (dh, dw) in enumerate(zip(tap_dh, tap_dw)):
        acc += _gather(x, dh, dw, OH, OW, stride) @ wk[:, t * Cin:(t + 1) * Cin].t()
    if epi_scale is not None:
        acc = acc * epi_scale + epi_shift
    if act == 1:
        acc = acc.relu()
    elif act == 2:
        acc = torch.where(acc > 0, acc, acc * slope)
    y[:, oph::osh, opw::osw][:, :OH, :OW, y_coff:y_coff + Cout] = acc
    return y


def conv_wgrad(x, dy, tap_dh, tap_dw, Cout, OH, OW, stride=1, dy_coff=0, force_plan=-1, plan_out=None, real=None, **_):
    N, H, W, Cin = x.shape
    dw_ = torch.zeros(len(tap_dh), Cout, Cin)
    d = dy[..., dy_coff:dy_coff + Cout].reshape(-1, Cout)
    for t, (dh, dwo) in enumerate(zip(tap_dh, tap_dw)):
        dw_[t] = d.t() @ _gather(x, dh, dwo, OH, OW, stride).reshape(-1, Cin)
    return dw_


def pack_taps(w, taps, k_padded, round_tf32=True):
    """sos_pack_taps semantics: out[r, t*k_padded + k] = w[r, k, taps[t][0], taps[t][1]], zero padded (no rounding here)."""
    R, K = w.shape[0], w.shape[1]
    out = torch.zeros(R, len(taps) * k_padded)
    for t, tap in enumerate(taps):
        if tap is not None:                                                # None: a zero tap (padding of a folded operand row)
            out[:, t * k_padded:t * k_padded + K] = w[:, :, tap[0], tap[1]]
    return out


def im2col_half(x, tap_dh, tap_dw, OH, OW, Kc):
    """sos_im2col_half: column 2 t + c = x[n, oh + dh_t, ow + dw_t, c] for the two real channels, zero outside / in the padding."""
    out = torch.zeros(x.shape[0], OH, OW, Kc, dtype=torch.float16)
    for t, (dh, dw) in enumerate(zip(tap_dh, tap_dw)):
        out[..., 2 * t:2 * t + 2] = _gather(x[..., :2].float(), dh, dw, OH, OW, 1).half()
    return out


# ------------------------------------------------------------------------------------------------ half-path kernels (CPU semantics)
# Documented semantics of the half-operand entry points (include/sos_b200.h), enough to drive layers.ConvBNActH / TapConvH /
# PadCatH / ToHalf on the CPU: the host-side wiring (handles, channel padding, scale plumbing, autograd returns) is what is tested.
def _f(t):
    return t.float() if t.dtype == torch.float16 else t


def conv_tc_any(x, wk, tap_dh, tap_dw, Cout, OH, OW, stride=1, y=None, y_coff=0, lattice=(1, 1, 0, 0), epi_scale=None, epi_shift=None,
                act=0, slope=None, force_plan=-1, plan_out=None, k_real=None, tag=None, want_stats=False, y_half=False, out_scale=None, bnr=None):
    """sos_conv2d_tc for fp32 or half operands (fp32 accumulate); optional output scale, half output, BatchNorm partial sums."""
    xf, wf = _f(x), _f(wk)
    N = x.shape[0]
    osh, osw, oph, opw = lattice
    acc = conv_tc(xf, wf, tap_dh, tap_dw, Cout, OH, OW, stride)[..., :Cout]           # dense (N, OH, OW, Cout) of THIS call
    if out_scale is not None:
        acc = acc * out_scale
    if epi_scale is not None:
        acc = acc * epi_scale[:Cout] + epi_shift[:Cout]
    a = act & 15
    if a == 1:
        acc = acc.relu()
    elif a == 2:
        acc = torch.where(acc > 0, acc, acc * slope)
    if y is None:
        y = torch.zeros(N, OH * osh, OW * osw, (Cout + 7) // 8 * 8, dtype=torch.float16 if y_half else torch.float32)
    y[:, oph::osh, opw::osw][:, :OH, :OW, y_coff:y_coff + Cout] = acc.to(y.dtype)
    if bnr is not None:
        return y, None                                 # (the emulation has no fused reduction: the caller runs pass 1 itself)
    if want_stats:
        flat = torch.nn.functional.pad(acc, (0, y.shape[3] - Cout)).reshape(-1, y.shape[3])
        return y, torch.stack([flat.sum(0), (flat * flat).sum(0)])[None]
    return y


def conv_wgrad_any(x, dy, tap_dh, tap_dw, Cout, OH, OW, stride=1, dy_coff=0, force_plan=-1, plan_out=None, real=None, out_scale=None,
                   workspace=False):
    dw = conv_wgrad(_f(x), _f(dy), tap_dh, tap_dw, Cout, OH, OW, stride, dy_coff)
    return dw * out_scale if out_scale is not None else dw


def pack_taps_half(w, taps, k_padded):
    return pack_taps(w, taps, k_padded).half()


def to_half(x, cd=None, scaled=False, out=None):
    cs = x.shape[-1]
    cd = cs if cd is None else cd
    s = 1.0
    scal = None
    if scaled:
        ms = float((x.double() ** 2).mean())
        e = 0 if ms <= 0 else max(-100, min(100, round(-0.5 * torch.log2(torch.tensor(ms)).item())))
        s = 2.0 ** e
        scal = torch.tensor([s, 1.0 / s, float((x.double() ** 2).sum())], dtype=torch.float32)
    h = torch.nn.functional.pad(x * s, (0, cd - cs)).clamp(-65504, 65504).half()
    if out is not None:
        out.copy_(h)
        h = out
    return (h, scal) if scaled else h


def bn_finalize_partial(conv_partial, rows, gamma, beta, running_mean, running_var, eps, momentum):
    tot = conv_partial.double().sum(0)
    mean = tot[0] / rows
    var = (tot[1] / rows - mean * mean).clamp_min(0)
    invstd = 1.0 / torch.sqrt(var + eps)
    scale = gamma.double() * invstd
    running_mean.mul_(1 - momentum).add_(momentum * mean.float())
    running_var.mul_(1 - momentum).add_(momentum * (var * rows / max(rows - 1, 1)).float())
    return torch.stack([mean, invstd, scale, beta.double() - mean * scale]).float()


def _act(pre, act, slope):
    a = act & 15
    if a == 1:
        return pre.relu()
    if a == 2:
        return torch.where(pre > 0, pre, pre * slope)
    return pre


def bn_act_apply(y, stats, act, slope, half):
    from sos_b200 import ops
    y = _f(y)                                             # (the raw conv output may be stored as half)
    z = _act(y * stats[2] + stats[3], act, slope)
    if not half:
        return z
    h = ops.new_half(y.shape, y.device)
    ops.hv(h).copy_(z.half())
    return h


def bn_train_backward_half(dz, y, stats, act, slope, grad_into=None, dz_inv=None, pre_partial=None):
    """dz / y may be stored as half; a half dz still carries the power-of-two scale of the layer above (inverse = dz_inv): the
    kernels work on the stored values, scale the parameter gradients by dz_inv and publish the COMPOSED scale of dy."""
    dz, y = _f(dz), _f(y)
    inv_in = float(dz_inv) if dz_inv is not None else 1.0
    mean, invstd, scale, shift = stats
    pre = y * scale + shift
    a = act & 15
    dpre = dz.clone()
    dslope = None
    if a == 1:
        dpre = torch.where(pre > 0, dz, torch.zeros_like(dz))
    elif a == 2:
        dslope = (dz * pre)[pre <= 0].sum().reshape(1)
        dpre = torch.where(pre > 0, dz, dz * slope)
    xhat = (y - mean) * invstd
    flat = lambda t: t.reshape(-1, t.shape[-1])
    dbeta, dgamma = flat(dpre).sum(0), flat(dpre * xhat).sum(0)
    n = flat(dpre).shape[0]
    dy = scale * (dpre - dbeta / n - xhat * (dgamma / n))
    dyh, scal = to_half(dy, scaled=True)
    scal = torch.tensor([float(scal[0]) / inv_in, float(scal[1]) * inv_in, float(scal[2])], dtype=torch.float32)
    dgamma, dbeta = dgamma * inv_in, dbeta * inv_in
    if dslope is not None:
        dslope = dslope * inv_in
    if grad_into is not None:
        gg, gb, gs = grad_into
        gg += dgamma[:gg.numel()]
        gb += dbeta[:gb.numel()]
        if gs is not None and dslope is not None:
            gs += dslope
        return dyh, None, None, None, scal
    return dyh, dgamma, dbeta, dslope, scal


def copy_view(src, sview, dst, dview, batch, channels, accumulate=False):
    """sos_copy_view: dst window = nearest-resized src window, `channels` fp32 elements per pixel (both as (N, Hp, Wp, ld) buffers)."""
    sH, sW, sHp, sWp, sph, spw, sld, scoff = list(sview)
    dH, dW, dHp, dWp, dph, dpw, dld, dcoff = list(dview)
    s = src.reshape(batch, sHp, sWp, sld)[:, sph:sph + sH, spw:spw + sW, scoff:scoff + channels]
    hi = torch.clamp((torch.arange(dH) * (sH / dH)).floor().long(), max=sH - 1) if sH != dH else torch.arange(dH)
    wi = torch.clamp((torch.arange(dW) * (sW / dW)).floor().long(), max=sW - 1) if sW != dW else torch.arange(dW)
    win = dst.reshape(batch, dHp, dWp, dld)[:, dph:dph + dH, dpw:dpw + dW, dcoff:dcoff + channels]
    val = s[:, hi][:, :, wi]
    win.copy_(win + val if accumulate else val)


def reflect_fill(buf, H, W, pad):
    inner = buf[:, pad:pad + H, pad:pad + W].permute(0, 3, 1, 2)
    buf.copy_(torch.nn.functional.pad(inner, (pad, pad, pad, pad), mode="reflect").permute(0, 2, 3, 1))


def reflect_fold(gbuf, H, W, pad):
    """sos_reflect_fold: adjoint of the reflect fill, in place (the interior collects the border gradients mirrored onto it)."""
    with torch.enable_grad():
        inner = torch.zeros(gbuf.shape[0], gbuf.shape[3], H, W, requires_grad=True)
        torch.nn.functional.pad(inner, (pad, pad, pad, pad), mode="reflect").backward(gbuf.detach().permute(0, 3, 1, 2))
    gbuf[:, pad:pad + H, pad:pad + W] = inner.grad.permute(0, 2, 3, 1)


def copy_view_fold(gpad, pview, dst, dview, batch, channels):
    """sos_copy_view_fold: dst window = folded interior of the reflect-padded gradient map (which is left unchanged)."""
    H, W, Hp, Wp, ph, pw, ld, coff = list(pview)
    g = gpad.reshape(batch, Hp, Wp, ld)[..., coff:coff + channels].clone()
    reflect_fold(g, H, W, ph)
    dH, dW, dHp, dWp, dph, dpw, dld, dcoff = list(dview)
    dst.reshape(batch, dHp, dWp, dld)[:, dph:dph + dH, dpw:dpw + dW, dcoff:dcoff + channels] = g[:, ph:ph + H, pw:pw + W]


def copy_view_backward(gdst, dview, gsrc, sview, batch, channels):
    """sos_copy_view_backward: gsrc window += gdst window through the nearest map of copy_view."""
    sH, sW, sHp, sWp, sph, spw, sld, scoff = list(sview)
    dH, dW, dHp, dWp, dph, dpw, dld, dcoff = list(dview)
    g = gdst.reshape(batch, dHp, dWp, dld)[:, dph:dph + dH, dpw:dpw + dW, dcoff:dcoff + channels]
    hi = torch.clamp((torch.arange(dH) * (sH / dH)).floor().long(), max=sH - 1) if sH != dH else torch.arange(dH)
    wi = torch.clamp((torch.arange(dW) * (sW / dW)).floor().long(), max=sW - 1) if sW != dW else torch.arange(dW)
    acc = torch.zeros(batch, sH, sW, channels)
    tmp = torch.zeros(batch, sH, dW, channels).index_add_(1, hi, g)
    acc.index_add_(2, wi, tmp)
    win = gsrc.reshape(batch, sHp, sWp, sld)[:, sph:sph + sH, spw:spw + sW, scoff:scoff + channels]
    win.copy_(win + acc)
