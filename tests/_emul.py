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


def conv_tc(x, wk, tap_dh, tap_dw, Cout, OH, OW, stride=1, y=None, y_coff=0, lattice=(1, 1, 0, 0), epi_scale=None, epi_shift=None,
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
    for t, (a, b) in enumerate(taps):
        out[:, t * k_padded:t * k_padded + K] = w[:, :, a, b]
    return out
