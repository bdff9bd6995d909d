"""Autograd building blocks over the sm_100a kernels (NHWC fp32 activations).

Each Function's forward and backward enqueue kernels of libsos_b200.so through `ops`; PyTorch autograd only
records the graph.  Reference semantics:
  TapConv       nn.Conv2d / nn.ConvTranspose2d (bias=False) of M1/networks.py:36, M2/networks.py:36,105-106,130-131
  BNAct         nn.BatchNorm2d + ReLU / PReLU   of the same blocks
  PadCat        torch.cat + F.interpolate(nearest) + nn.ReflectionPad2d (M2/networks.py:104,129,198-204)
  FeatToSeq     view/interpolate/permute before the LSTMs (M1/networks.py:131-135, M2/networks.py:83-86)
  BiLSTM        nn.LSTM(bidirectional=True) (M1/networks.py:95, M2/networks.py:64)
"""
import os

import torch

from . import ops


def _round8(c):
    return (c + 7) // 8 * 8


# ----------------------------------------------------------------------------------------------- arithmetic mode
# Default: every convolution is ONE tensor-core pass over TF32-rounded operands (the contract cuDNN applies to the
# reference's nn.Conv2d on Ampere+ GPUs).  Precise mode (verification / fp32-parity runs, 3x the tensor work): operands are
# split x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and a convolution is hi*hi + lo*hi + hi*lo (error ~2^-21, i.e.
# fp32-grade), forward, data gradient and weight gradient alike; producers then do not round.
_PRECISE = False


def set_precise(flag):
    global _PRECISE
    old, _PRECISE = _PRECISE, bool(flag)
    return old


def precise():
    return _PRECISE


# Operand storage of the default (one-pass) mode: IEEE half (True; tcgen05 kind::f16 -- the same 11-bit significand as TF32 at
# half the bytes and twice the tensor rate; gradient operands carry a per-tensor power-of-two scale) or TF32 in fp32 storage
# (False; kind::tf32, the first implementation, kept for A/B measurements).  Precise mode always runs on the TF32 kernels.
_HALF = True


def set_half(flag):
    global _HALF
    old, _HALF = _HALF, bool(flag)
    return old


def half_mode():
    return _HALF and not _PRECISE


# Weight gradients are off the backward critical path (only the optimiser needs them): inside `async_wgrad()` TapConv.backward
# enqueues them on a side stream and adds them straight into the parameter's .grad (a view of the agent's flat gradient buffer),
# so the tensor-bound wgrad kernels overlap the HBM-bound BatchNorm passes of the following layers.  The caller joins the
# streams (`join_wgrad()`) before it reads any gradient.
_ASYNC_WGRAD = False
_SIDE = {}                         # main stream -> its weight-gradient side stream (concurrent graph branches get one each)
_DIRECT_GRADS = os.environ.get("SOS_DIRECT_GRADS", "1") != "0"      # A/B switch: parameter gradients added straight into .grad


def _side_stream():
    key = torch.cuda.current_stream().cuda_stream
    side = _SIDE.get(key)
    if side is None:
        side = _SIDE[key] = torch.cuda.Stream()
    return side


class async_wgrad(object):
    def __enter__(self):
        global _ASYNC_WGRAD
        self.old, _ASYNC_WGRAD = _ASYNC_WGRAD, True

    def __exit__(self, *a):
        global _ASYNC_WGRAD
        _ASYNC_WGRAD = self.old
        join_wgrad()


def join_wgrad():
    side = _SIDE.get(torch.cuda.current_stream().cuda_stream)
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)


def _split(t):
    hi = ops.round_tf32_(t.detach().clone().contiguous())
    lo = ops.round_tf32_((t.detach() - hi).contiguous())
    return hi, lo


# ----------------------------------------------------------------------------------------------- geometry
class ConvGeom:
    """Tap lists of one convolution in the three roles (forward, data gradient, weight gradient)."""

    def __init__(self, kind, kh, kw, dh=1, dw=1, stride=1, round_dy=False):
        assert kind in ("zero", "valid", "convT")
        self.kind, self.kh, self.kw, self.dh, self.dw, self.stride = kind, kh, kw, dh, dw, stride
        # True when the output gradient does not come from a BatchNorm backward kernel (which rounds it to TF32 itself)
        self.round_dy = round_dy
        if kind == "convT":
            assert (kh, kw, stride) == (3, 3, 2)
        taps = [(a, b) for a in range(kh) for b in range(kw)]
        self.taps = taps
        if kind == "zero":
            self.off = [((a - (kh - 1) // 2) * dh, (b - (kw - 1) // 2) * dw) for a, b in taps]
        elif kind == "valid":
            self.off = [(a * dh, b * dw) for a, b in taps]
        else:
            self.off = [(a - 1, b - 1) for a, b in taps]        # offsets of dy relative to 2*iy (used by dgrad / wgrad)

    def out_size(self, H, W):
        if self.kind == "zero":
            return H, W
        if self.kind == "valid":
            return (H - (self.kh - 1) * self.dh - 1) // self.stride + 1, (W - (self.kw - 1) * self.dw - 1) // self.stride + 1
        return 2 * H, 2 * W


# Packed GEMM operands of the weights are cached per (weight storage, view, tap list): inference packs every weight ONCE, not once
# per call.  An entry is valid while the weight's version counter and the global weight epoch are unchanged -- the optimiser
# updates parameters through the flat buffer behind autograd's back, so FlatAdam / load_ckpt bump the epoch (`weights_changed`).
# The entry holds the weight's storage, so its address cannot be recycled for another tensor while the entry lives.
# Training re-packs every weight once per optimiser step (133 packs: forward and data-gradient layouts): `pack_all()` refreshes ALL
# cached entries IN PLACE with one launch (sos_pack_taps_half_multi over a device table of descriptors); the trainers call it at the
# start of a step, and it is the first node of a captured step graph.  While a CUDA graph is being captured an entry is served only
# if that capture contains a `pack_all()`; anything else is packed by its own (captured) launch, so a replay never reads stale weights.
_WEIGHT_EPOCH = 0
_PACKS = {}                       # key -> [ver, wk, storage, weight view, descriptor]
_PACK_TABLE = [None, 0]           # device copy of the descriptors, number of entries it holds
_CAPTURE_EPOCH = -1               # weight epoch of the capture whose graph holds a pack_all()
_BATCH_PACK = os.environ.get("SOS_BATCH_PACK", "1") != "0"        # A/B switch


def weights_changed():
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def _pack_key(w, taps, cin_p):
    return (w.untyped_storage().data_ptr(), w.storage_offset(), tuple(w.shape), tuple(w.stride()), tuple(taps), cin_p)


def pack_all():
    """Bring every cached half weight operand up to date with ONE launch (no-op when nothing changed since the last call)."""
    global _CAPTURE_EPOCH
    if not (_BATCH_PACK and _PACKS):
        return
    capturing = torch.cuda.is_current_stream_capturing()
    ents = list(_PACKS.values())
    if not capturing and all(e[0] == (e[3]._version, _WEIGHT_EPOCH) for e in ents):
        return
    if _PACK_TABLE[0] is None or _PACK_TABLE[1] != len(ents):
        if capturing:
            return                                                  # (a table upload cannot be captured: the packs stay individual launches)
        _PACK_TABLE[0] = ops.pack_table([e[4] for e in ents], ents[0][1].device)
        _PACK_TABLE[1] = len(ents)
    ops.pack_taps_half_multi(_PACK_TABLE[0], len(ents))
    for e in ents:
        e[0] = (e[3]._version, _WEIGHT_EPOCH)
    if capturing:
        _CAPTURE_EPOCH = _WEIGHT_EPOCH


def _pack_fwd(w, taps, cin_p, half=False):
    """(Cout, Cin, kh, kw) view -> (Cout, len(taps)*cin_p), k = t*cin_p + ci, TF32-rounded fp32 or half (one gather kernel)."""
    if not half:
        return ops.pack_taps(w.detach(), taps, cin_p, round_tf32=True)
    key = _pack_key(w, taps, cin_p)
    ent = _PACKS.get(key)
    if w.is_cuda and torch.cuda.is_current_stream_capturing():
        if ent is not None and _CAPTURE_EPOCH == _WEIGHT_EPOCH:
            return ent[1]                                           # refreshed by the pack_all() node of this step's graph
        return ops.pack_taps_half(w.detach(), taps, cin_p)
    ver = (w._version, _WEIGHT_EPOCH)
    if ent is not None:
        if ent[0] != ver:                                           # stale: refresh in place (the entry's buffer is in the batch table)
            ops.pack_taps_half(w.detach(), taps, cin_p, out=ent[1])
            ent[0] = ver
        return ent[1]
    wk = ops.pack_taps_half(w.detach(), taps, cin_p)
    _PACKS[key] = [ver, wk, w.untyped_storage(), w.detach(), ops.pack_desc(w.detach(), taps, cin_p, wk)]
    return wk


# Two-channel inputs (the real / imaginary spectrogram planes every network starts from): a k x k convolution over 2 of 8 / 16 stored
# channels is 7-25 tap GEMMs over 16- / 32-byte pixel rows -- thousands of tiny TMA requests per tile, the input re-fetched per tap
# group (profiles/r02c_*_ncu.txt: the main loop starves the epilogue).  Folding the taps into the channel axis first
# (ops.im2col_half: K = 2 * ntaps real columns in one 32- / 64- / 128-byte row per pixel) turns a pass into a ONE-tap GEMM.
# Measured (scripts/bench_conv.py, batch 32): the weight gradient gains (2 -> 64 5x5: 0.274 -> 0.228 ms, 2 -> 96 1x7: 0.146 -> 0.126),
# the forward pass does not (0.166 -> 0.262, 0.186 -> 0.214: a one-tap 64 -> 64 GEMM is bound by its epilogue and the im2col pass
# comes on top), so only the weight gradient takes the folded form by default (_FOLD_FWD: SOS_FOLD_TAPS=2 folds the forward pass
# too); the data gradient (needed only where the input is another network's output) always keeps the tap form.
_FOLD_TAPS = os.environ.get("SOS_FOLD_TAPS", "1") != "0"           # A/B switch
_FOLD_FWD = os.environ.get("SOS_FOLD_TAPS", "1") == "2"


def _fold_kc(x, w, g):
    """Folded row width (halves) when this half-mode convolution qualifies for the folded form, else 0."""
    if not (_FOLD_TAPS and x.dtype == torch.float16 and g.kind != "convT" and g.stride == 1 and w.shape[1] == 2 and 1 < len(g.taps) <= 32):
        return 0
    k = 2 * len(g.taps)
    return 16 if k <= 16 else (32 if k <= 32 else 64)


def _fold_taps(g, kc):
    return list(g.taps) + [None] * (kc // 2 - len(g.taps))


def _conv_forward(x, w, g, epi=None, want_stats=False, y_half=False, y_out=None):
    """x NHWC (N,H,W,Cin_p); w PyTorch layout.  Returns y NHWC (N,OH,OW,round8(Cout)); with want_stats also the BatchNorm
    partial sums (G, 2, C) of y computed by the epilogue (the four sub-pixel launches of a transposed conv stack theirs)."""
    N, H, W, cin_p = x.shape
    OH, OW = g.out_size(H, W)
    kw = dict(epi_scale=epi[0], epi_shift=epi[1], act=epi[2], slope=epi[3]) if epi else {}
    half = x.dtype == torch.float16
    kw["y_half"] = y_half
    if g.kind != "convT":
        Cout = w.shape[0]
        kc = _fold_kc(x, w, g) if _FOLD_FWD else 0
        if kc:
            xcol = ops.im2col_half(x, [o[0] for o in g.off], [o[1] for o in g.off], OH, OW, kc)
            wk = _pack_fwd(w, _fold_taps(g, kc), 2, True)                      # (Cout, kc): column 2 t + c
            return ops.conv_tc(xcol, wk, [0], [0], Cout, OH, OW, 1, k_real=2 * len(g.taps), want_stats=want_stats, y=y_out, **kw)
        wk = _pack_fwd(w, g.taps, cin_p, half)
        return ops.conv_tc(x, wk, [o[0] for o in g.off], [o[1] for o in g.off], Cout, OH, OW, g.stride, k_real=w.shape[1],
                           want_stats=want_stats, y=y_out, **kw)
    # ConvTranspose2d(k3, s2, p1, output_padding=1): oy = 2*iy - 1 + ky  ->  four sub-pixel convolutions
    Cout = w.shape[1]
    wc = w.permute(1, 0, 2, 3)                                             # (Cout, Cin, ky, kx)
    Cy = _round8(Cout)
    y = y_out if y_out is not None else (torch.zeros if Cy != Cout else torch.empty)(N, OH, OW, Cy, device=x.device,
                                                                                     dtype=torch.float16 if y_half else torch.float32)
    ph_taps = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}                          # phase -> [(k, input offset)]
    parts = []
    for py in (0, 1):
        for px in (0, 1):
            taps = [(ky, kx) for ky, _ in ph_taps[py] for kx, _ in ph_taps[px]]
            offs = [(oy, ox) for _, oy in ph_taps[py] for _, ox in ph_taps[px]]
            wk = _pack_fwd(wc, taps, cin_p, half)
            r = ops.conv_tc(x, wk, [o[0] for o in offs], [o[1] for o in offs], Cout, H, W, 1, y=y, lattice=(2, 2, py, px), k_real=w.shape[0],
                            want_stats=want_stats, **kw)
            if want_stats:
                parts.append(r[1])
    return (y, torch.cat(parts)) if want_stats else y


def _conv_dgrad(dy, w, g, x_shape, out_scale=None):
    dx = _conv_dgrad_raw(dy, w, g, x_shape, out_scale)
    if dx.shape[3] != x_shape[3]:                        # x carries more zero channels than round8(Cin) (2 -> 16 in half mode)
        dx = torch.nn.functional.pad(dx, (0, x_shape[3] - dx.shape[3]))
    return dx


def _conv_dgrad_raw(dy, w, g, x_shape, out_scale=None, y_half=False, bnr=None):
    """Gradient w.r.t. the (possibly reflect-padded) NHWC input buffer (fp32).  dy: fp32 (TF32) or half with its inverse
    scale `out_scale`.  y_half (stride-1 convolutions only): the result is stored as half -- pass out_scale=None then, so that it
    keeps dy's power-of-two scale (see EncoderChainH)."""
    N, H, W, cin_p = x_shape
    cout_p = dy.shape[3]
    half = dy.dtype == torch.float16
    assert not y_half or (g.kind != "convT" and g.stride == 1)
    if g.kind == "convT":
        # dx[iy] = sum_k dy[2*iy - 1 + ky] w[ci][co][ky]: a stride-2 tap conv over dy
        Cin = w.shape[0]
        wk = _pack_fwd(w, g.taps, cout_p, half)                            # rows ci, k = t*cout_p + co
        return ops.conv_tc(dy, wk, [o[0] for o in g.off], [o[1] for o in g.off], Cin, H, W, 2, k_real=w.shape[1], tag="conv_dgrad",
                           out_scale=out_scale)
    Cin = w.shape[1]
    wt = w.permute(1, 0, 2, 3)                                             # (Cin, Cout, kh, kw)
    if g.stride == 1:
        wk = _pack_fwd(wt, g.taps, cout_p, half)
        return ops.conv_tc(dy, wk, [-o[0] for o in g.off], [-o[1] for o in g.off], Cin, H, W, 1, k_real=w.shape[0], tag="conv_dgrad",
                           out_scale=out_scale, y_half=y_half, bnr=bnr)
    # stride 2: four output phases of the input lattice
    dx = torch.empty(N, H, W, _round8(Cin), device=dy.device, dtype=torch.float32)
    assert _round8(Cin) == Cin
    for py in (0, 1):
        for px in (0, 1):
            sel = [(t, o) for t, o in zip(g.taps, g.off) if (py - o[0]) % 2 == 0 and (px - o[1]) % 2 == 0]
            oh, ow = (H - py + 1) // 2, (W - px + 1) // 2
            if not sel:
                dx[:, py::2, px::2].zero_()
                continue
            wk = _pack_fwd(wt, [t for t, _ in sel], cout_p, half)
            ops.conv_tc(dy, wk, [(py - o[0]) // 2 for _, o in sel], [(px - o[1]) // 2 for _, o in sel], Cin, oh, ow, 1, y=dx,
                        lattice=(2, 2, py, px), k_real=w.shape[0], tag="conv_dgrad", out_scale=out_scale)
    return dx


def _conv_wgrad_raw(x, dy, w, g, out_scale=None, workspace=False):
    """The weight-gradient kernel's tap-major buffer (ntaps, RP, CP): rows / columns are w's first two axes, zero padded."""
    if g.kind == "convT":
        Cin, Cout = w.shape[0], w.shape[1]
        # roles swapped: "input" = dy (2H x 2W), "output grad" = x (H x W), stride 2  ->  (ntaps, Cin_p, Cout_p)
        return ops.conv_wgrad(dy, x, [o[0] for o in g.off], [o[1] for o in g.off], x.shape[3], x.shape[1], x.shape[2], 2, real=(Cout, Cin),
                              out_scale=out_scale, workspace=workspace)
    Cout, Cin = w.shape[0], w.shape[1]
    OH, OW = dy.shape[1], dy.shape[2]
    kc = _fold_kc(x, w, g)
    if kc:                                                                     # folded: (1, Cout_p, kc), column 2 t + c
        xcol = ops.im2col_half(x, [o[0] for o in g.off], [o[1] for o in g.off], OH, OW, kc)
        return ops.conv_wgrad(xcol, dy, [0], [0], dy.shape[3], OH, OW, 1, real=(2 * len(g.taps), Cout), out_scale=out_scale, workspace=workspace)
    return ops.conv_wgrad(x, dy, [o[0] for o in g.off], [o[1] for o in g.off], dy.shape[3], OH, OW, g.stride, real=(Cin, Cout),
                          out_scale=out_scale, workspace=workspace)


def _unfold_wgrad(dwt, w, g):
    """The folded weight-gradient buffer (1, Cout_p, kc) as a view in w's own layout (Cout, 2, kh, kw)."""
    R, nt = w.shape[0], len(g.taps)
    return dwt[0, :R, :2 * nt].reshape(R, nt, 2).permute(0, 2, 1).reshape(R, 2, g.kh, g.kw)


def _conv_wgrad(x, dy, w, g, out_scale=None):
    """Returns the gradient in w's own layout."""
    dwt = _conv_wgrad_raw(x, dy, w, g, out_scale)
    if _fold_kc(x, w, g):
        return _unfold_wgrad(dwt, w, g).contiguous()
    R, Cc = w.shape[0], w.shape[1]
    return dwt[:, :R, :Cc].permute(1, 2, 0).reshape(R, Cc, g.kh, g.kw).contiguous()


class TapConv(torch.autograd.Function):
    """forward(x, w, geometry, want_stats=False) -> y, or (y, partial) with the BatchNorm partial sums of y from the epilogue."""

    @staticmethod
    def forward(ctx, x, w, g, want_stats=False):
        ctx.g = g
        ctx.save_for_backward(x, w)
        ctx.precise = _PRECISE
        if _PRECISE:
            (xh, xl), (wh, wl) = _split(x), _split(w)
            y = _conv_forward(xh, wh, g) + _conv_forward(xl, wh, g) + _conv_forward(xh, wl, g)
            return (y, torch.empty(0, device=y.device)) if want_stats else y       # precise mode: statistics from the summed y
        if want_stats:
            y, partial = _conv_forward(x, w, g, want_stats=True)
            ctx.mark_non_differentiable(partial)
            return y, partial
        return _conv_forward(x, w, g)

    @staticmethod
    def backward(ctx, dy, *unused):
        x, w = ctx.saved_tensors
        g = ctx.g
        dy = dy.contiguous()
        if ctx.precise:
            (xh, xl), (wh, wl), (dh, dl) = _split(x), _split(w), _split(dy)
            dx = dw = None
            if ctx.needs_input_grad[0]:
                dx = _conv_dgrad(dh, wh, g, x.shape) + _conv_dgrad(dl, wh, g, x.shape) + _conv_dgrad(dh, wl, g, x.shape)
            if ctx.needs_input_grad[1]:
                dw = _conv_wgrad(xh, dh, w, g) + _conv_wgrad(xl, dh, w, g) + _conv_wgrad(xh, dl, w, g)
            return dx, dw, None, None
        if g.round_dy:
            dy = ops.round_tf32_(dy.clone())
        dw = None
        if ctx.needs_input_grad[1] and _ASYNC_WGRAD and w.grad is not None:
            side, main = _side_stream(), torch.cuda.current_stream()
            side.wait_stream(main)                                     # x, dy (and the zeroed .grad) are ready
            with torch.cuda.stream(side):
                w.grad.add_(_conv_wgrad(x, dy, w, g))
            x.record_stream(side)
            dy.record_stream(side)
        elif ctx.needs_input_grad[1]:
            dw = _conv_wgrad(x, dy, w, g)
        dx = _conv_dgrad(dy, w, g, x.shape) if ctx.needs_input_grad[0] else None
        return dx, dw, None, None


def conv_fused_eval(x, w, g, scale, shift, act, slope, round_out=True):
    """Inference path: BN (running stats) + activation folded into the GEMM epilogue.  Half mode: x is a half-map handle and
    the epilogue stores half (round_out) or fp32 (the map feeds the LSTM projection)."""
    if ops.is_half_handle(x):
        if not round_out:
            return _conv_forward(ops.hv(x), w, g, epi=(scale, shift, act, slope))
        Cout = w.shape[1] if g.kind == "convT" else w.shape[0]
        assert Cout % 8 == 0
        OH, OW = g.out_size(x.shape[1], x.shape[2])
        z = ops.new_half((x.shape[0], OH, OW, Cout), x.device)
        _conv_forward(ops.hv(x), w, g, epi=(scale, shift, act, slope), y_half=True, y_out=ops.hv(z))
        return z
    return _conv_forward(x, w, g, epi=(scale, shift, act | (ops.ACT_ROUND_TF32 if round_out else 0), slope))


def _wgrad_into(w, x, dy, g, out_scale):
    """Weight gradient: on the side stream straight into w.grad inside async_wgrad(), else returned."""
    if _ASYNC_WGRAD and w.grad is not None:
        side, main = _side_stream(), torch.cuda.current_stream()
        side.wait_stream(main)                                     # x, dy, the scale (and the zeroed .grad) are ready
        with torch.cuda.stream(side):
            if _DIRECT_GRADS and w.grad.is_contiguous() and _fold_kc(x, w, g):
                dwt = _conv_wgrad_raw(x, dy, w, g, out_scale, workspace=_WGRAD_WS)
                w.grad.add_(_unfold_wgrad(dwt, w, g))
                if _WGRAD_WS:
                    dwt.zero_()                                   # (the shared workspace is handed back empty)
            elif _DIRECT_GRADS and w.grad.is_contiguous():
                # one pass: un-pad, re-layout, add -- out of the side stream's shared workspace, which it leaves zeroed (no fill launch)
                ops.accumulate_wgrad(_conv_wgrad_raw(x, dy, w, g, out_scale, workspace=_WGRAD_WS), w.grad, clear=_WGRAD_WS)
            else:
                w.grad.add_(_conv_wgrad(x, dy, w, g, out_scale))
        x.record_stream(side)
        dy.record_stream(side)
        if out_scale is not None:
            out_scale.record_stream(side)
        return None
    return _conv_wgrad(x, dy, w, g, out_scale)


# BatchNorm-backward reduction inside the data-gradient epilogue of the row-streaming kernel (sos_conv_args::bnr_*).  OFF by default:
# measured (scripts/bench_bnr.py, 48 -> 48 5x5 at batch 32) the data gradient goes from 0.168 to 0.262 ms -- the three-quantity warp
# transpose-reduce makes its epilogue the bottleneck -- while the BatchNorm backward drops from 0.190 to 0.102 ms: a wash per layer and
# in the step (79.3-79.9 ms either way).  SOS_FUSE_BNR=1 turns it on.
_FUSE_BNR = os.environ.get("SOS_FUSE_BNR", "0") == "1"
_WGRAD_WS = os.environ.get("SOS_WGRAD_WS", "1") != "0"           # A/B switch: weight gradients through the shared zeroed workspace
_Y_HALF = os.environ.get("SOS_Y_HALF", "1") != "0"                 # A/B switch: raw conv outputs (BatchNorm inputs) stored as half


# nn.BatchNorm2d.num_batches_tracked += 1 in training mode (M2/networks.py:37; nobody reads it, but it is part of the state_dict):
# the blocks note their counters here and the networks bump them all with ONE multi-tensor launch per forward instead of 60.
_BN_PENDING = []


def note_bn_step(bn):
    _BN_PENDING.append(bn.num_batches_tracked)


def flush_bn_steps():
    if _BN_PENDING:
        with torch.no_grad():
            torch._foreach_add_(_BN_PENDING, 1)
        _BN_PENDING.clear()


def _bn_pad(gamma, beta, rm, rv, Cp):
    """BatchNorm parameter vectors zero-padded to the map's channel count (gamma = beta = 0 -> z = 0 in the padded channels)."""
    Cn = gamma.numel()
    gm, bt = gamma.detach(), beta.detach()
    if Cp != Cn:
        pad = (0, Cp - Cn)
        gm, bt, rm = (torch.nn.functional.pad(t, pad) for t in (gm, bt, rm))
        rv = torch.nn.functional.pad(rv, pad, value=1.0)
    return gm.contiguous(), bt.contiguous(), rm, rv


class EncoderChainH(torch.autograd.Function):
    """A whole encoder -- a chain of Conv2d(bias=False, zero 'same' padding, dilation) + BatchNorm2d(batch statistics) + ReLU blocks
    (`encoder_audio` M1/networks.py:91-93, `encoder_x` / `encoder_n` M2/networks.py:72-80) -- as ONE autograd node over half maps.

    Inside the node nothing travels in fp32: a block's raw conv output y and its activation z are half maps, and in backward the
    data gradient of block i + 1 is stored by its GEMM epilogue as half WITHOUT undoing the operand scale of dy_{i+1}; block i's
    BatchNorm backward (linear in dz) reads it together with that scale's inverse.  Per element and block: 4 B forward + 10 B
    backward of BatchNorm traffic instead of 6 + 18.  The last block (-> 8 / 4 channels, feeds the LSTM) keeps fp32 y and z.

    forward(x handle, geoms, eps, momentum, *[w, gamma, beta, running_mean, running_var] per block) -> z_last fp32 dense."""

    @staticmethod
    def forward(ctx, x, geoms, eps, momentum, *params):
        n = len(geoms)
        assert len(params) == 5 * n
        saved, cur = [], x
        for i in range(n):
            w, gamma, beta, rm, rv = params[5 * i:5 * i + 5]
            last = i == n - 1
            xh = ops.hv(cur)
            y, partial = _conv_forward(xh, w, geoms[i], want_stats=True, y_half=not last)
            Cp, Cn = y.shape[3], gamma.numel()
            gm, bt, rmp, rvp = _bn_pad(gamma, beta, rm, rv, Cp)
            stats = ops.bn_finalize_partial(partial, y.numel() // Cp, gm, bt, rmp, rvp, eps, momentum)
            if Cp != Cn:
                rm.copy_(rmp[:Cn])
                rv.copy_(rvp[:Cn])
            z = ops.bn_act_apply(y, stats, ops.ACT_RELU, None, half=not last)
            saved += [cur, y, stats]
            cur = z
        ctx.geoms, ctx.n = geoms, n
        ctx.save_for_backward(*saved, *[params[5 * i + k] for i in range(n) for k in range(3)])
        return cur

    @staticmethod
    def backward(ctx, dz):
        n, geoms = ctx.n, ctx.geoms
        t = ctx.saved_tensors
        acts, prm = t[:3 * n], t[3 * n:]
        grads = [None] * (5 * n)
        dz, dz_inv = dz.contiguous(), None
        dx = None
        pre = None                                    # pass 1 of this block's BatchNorm backward, when the GEMM above already did it
        for i in reversed(range(n)):
            x, y, stats = acts[3 * i:3 * i + 3]
            w, gamma, beta = prm[3 * i:3 * i + 3]
            g, Cn = geoms[i], gamma.numel()
            direct = (_ASYNC_WGRAD and _DIRECT_GRADS and gamma.grad is not None and beta.grad is not None and gamma.grad.is_contiguous()
                      and beta.grad.is_contiguous())
            dy, dgamma, dbeta, _, scal = ops.bn_train_backward_half(dz, y, stats, ops.ACT_RELU, None, grad_into=(gamma.grad, beta.grad, None) if direct else None,
                                                                    dz_inv=dz_inv, pre_partial=pre)
            pre = None
            inv = scal[1:2]
            xh = ops.hv(x)
            need_w = ctx.needs_input_grad[4 + 5 * i]
            dw = _wgrad_into(w, xh, dy, g, inv) if need_w else None
            grads[5 * i] = dw
            if not direct:
                grads[5 * i + 1], grads[5 * i + 2] = dgamma[:Cn], dbeta[:Cn]
            if i > 0:
                # data gradient stored as half, still carrying dy's scale (no out_scale): the next block's dz
                # (and, where the kernel can, the reduction pass of block i - 1's BatchNorm backward over the dz it is writing)
                y_below, stats_below = acts[3 * (i - 1) + 1], acts[3 * (i - 1) + 2]
                if _FUSE_BNR and y_below.dtype == torch.float16 and y_below.shape[:3] == xh.shape[:3]:
                    dz, pre = _conv_dgrad_raw(dy, w, g, xh.shape, None, y_half=True, bnr=(y_below, stats_below))
                else:
                    dz = _conv_dgrad_raw(dy, w, g, xh.shape, None, y_half=True)
                dz_inv = inv
            elif ctx.needs_input_grad[0]:
                dx = _conv_dgrad(dy, w, g, xh.shape, inv)
        return (dx, None, None, None, *grads)


class ConvBNActH(torch.autograd.Function):
    """Training-mode conv -> BatchNorm (batch statistics from the GEMM epilogue) -> ReLU / PReLU over HALF maps, as ONE autograd
    node: the gradient w.r.t. the conv output is a scaled half operand that never leaves this node.
    x: half-map handle; returns a half-map handle (or a dense fp32 map with out_f32)."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, slope, running_mean, running_var, eps, momentum, act, g, out_f32):
        xh = ops.hv(x)
        # raw conv output stored as half unless the block's output stays fp32 (statistics always come from the fp32 accumulators)
        y, partial = _conv_forward(xh, w, g, want_stats=True, y_half=not out_f32 and _Y_HALF)
        Cp, Cn = y.shape[3], gamma.numel()
        gm, bt, rm, rv = gamma.detach(), beta.detach(), running_mean, running_var
        if Cp != Cn:                                                      # padded channels: gamma = beta = 0 -> z = 0
            pad = (0, Cp - Cn)
            gm, bt, rm = (torch.nn.functional.pad(t, pad) for t in (gm, bt, rm))
            rv = torch.nn.functional.pad(rv, pad, value=1.0)
        stats = ops.bn_finalize_partial(partial, y.numel() // Cp, gm.contiguous(), bt.contiguous(), rm, rv, eps, momentum)
        if Cp != Cn:
            running_mean.copy_(rm[:Cn])
            running_var.copy_(rv[:Cn])
        z = ops.bn_act_apply(y, stats, act, slope, half=not out_f32)
        ctx.g, ctx.act, ctx.Cn = g, act, Cn
        ctx.save_for_backward(x, w, y, stats, slope if slope is not None else torch.empty(0), gamma, beta)
        return z

    @staticmethod
    def backward(ctx, dz):
        x, w, y, stats, slope, gamma, beta = ctx.saved_tensors
        slope = slope if slope.numel() else None
        g, Cn = ctx.g, ctx.Cn
        # inside async_wgrad() (the agents' update_network) the BatchNorm parameter gradients are added straight into .grad (views
        # of the flat gradient buffer) by the finalize kernel: no gradient tensors for autograd to accumulate
        direct = (_ASYNC_WGRAD and _DIRECT_GRADS and gamma.grad is not None and beta.grad is not None and (slope is None or slope.grad is not None)
                  and gamma.grad.is_contiguous() and beta.grad.is_contiguous())
        into = (gamma.grad, beta.grad, slope.grad if slope is not None else None) if direct else None
        dy, dgamma, dbeta, dslope, scal = ops.bn_train_backward_half(dz.contiguous(), y, stats, ctx.act, slope, grad_into=into)
        inv = scal[1:2]
        xh = ops.hv(x)
        dw = _wgrad_into(w, xh, dy, g, inv) if ctx.needs_input_grad[1] else None
        dx = _conv_dgrad(dy, w, g, xh.shape, inv) if ctx.needs_input_grad[0] else None
        if direct:
            return dx, dw, None, None, None, None, None, None, None, None, None, None
        return dx, dw, dgamma[:Cn], dbeta[:Cn], dslope, None, None, None, None, None, None, None


class TapConvH(torch.autograd.Function):
    """Convolution of a half map with an fp32 output and no normalisation after it (the last InpaintNet layer, and the
    eval-with-autograd path): forward(x handle, w, geometry) -> y fp32."""

    @staticmethod
    def forward(ctx, x, w, g):
        ctx.g = g
        ctx.save_for_backward(x, w)
        return _conv_forward(ops.hv(x), w, g)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        g = ctx.g
        dyh, scal = ops.to_half(dy.contiguous(), scaled=True)
        inv = scal[1:2]
        xh = ops.hv(x)
        dw = _wgrad_into(w, xh, dyh, g, inv) if ctx.needs_input_grad[1] else None
        dx = _conv_dgrad(dyh, w, g, xh.shape, inv) if ctx.needs_input_grad[0] else None
        return dx, dw, None


class ToHalf(torch.autograd.Function):
    """fp32 NHWC map -> half-map handle with the channels zero-padded to `cd` (straight-through gradient)."""

    @staticmethod
    def forward(ctx, x, cd):
        ctx.cs = x.shape[3]
        z = ops.new_half((*x.shape[:3], cd), x.device)
        ops.to_half(x.detach().contiguous(), cd, out=ops.hv(z))
        return z

    @staticmethod
    def backward(ctx, g):
        return (g if g.shape[3] == ctx.cs else g[..., :ctx.cs].contiguous()), None


def to_operand(x, cd=None):
    """Marks an fp32 NHWC map produced by a plain tensor expression as the input of a tensor-core GEMM."""
    if half_mode():
        return ToHalf.apply(x, cd or (x.shape[3] + 15) // 16 * 16)
    return RoundTF32.apply(x)


class RoundTF32(torch.autograd.Function):
    """Identity up to TF32 rounding of the value (straight-through gradient): marks a tensor that feeds a tensor-core GEMM
    but was produced by a plain tensor expression."""

    @staticmethod
    def forward(ctx, x):
        return x.detach().clone() if _PRECISE else ops.round_tf32_(x.detach().clone().contiguous())

    @staticmethod
    def backward(ctx, g):
        return g


# ----------------------------------------------------------------------------------------------- BatchNorm + act
class BNActTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, gamma, beta, slope, running_mean, running_var, eps, momentum, act, round_grad=True, conv_partial=None):
        z, stats = ops.bn_train_forward(y, gamma, beta, running_mean, running_var, eps, momentum, act, slope, conv_partial)
        ctx.act = (act & 15) | (ops.ACT_ROUND_TF32 if round_grad else 0)
        ctx.save_for_backward(y, stats, slope if slope is not None else torch.empty(0))
        return z

    @staticmethod
    def backward(ctx, dz):
        y, stats, slope = ctx.saved_tensors
        slope = slope if slope.numel() else None
        dy, dgamma, dbeta, dslope = ops.bn_train_backward(dz.contiguous(), y, stats, ctx.act, slope)
        return dy, dgamma, dbeta, dslope, None, None, None, None, None, None, None


def bn_act(y, bn, act, slope, training, round_out=True, round_grad=True, conv_partial=None):
    """y NHWC with C = round8(bn.num_features).  bn: nn.BatchNorm2d holding the reference-named parameters."""
    Cp, Cn = y.shape[3], bn.num_features
    if _PRECISE:
        round_out = round_grad = False
    gamma, beta, rm, rv = bn.weight, bn.bias, bn.running_mean, bn.running_var
    if Cp != Cn:                                                          # padded channels: gamma = beta = 0 -> z = 0
        pad = (0, Cp - Cn)
        gamma, beta = torch.nn.functional.pad(gamma, pad), torch.nn.functional.pad(beta, pad)
        rm_p, rv_p = torch.nn.functional.pad(rm, pad), torch.nn.functional.pad(rv, pad, value=1.0)
    else:
        rm_p, rv_p = rm, rv
    if training:
        if conv_partial is not None and conv_partial.numel() == 0:
            conv_partial = None
        z = BNActTrain.apply(y, gamma, beta, slope, rm_p, rv_p, bn.eps, bn.momentum, act | (ops.ACT_ROUND_TF32 if round_out else 0),
                             round_grad, conv_partial)
        if Cp != Cn:
            with torch.no_grad():
                rm.copy_(rm_p[:Cn])
                rv.copy_(rv_p[:Cn])
        note_bn_step(bn)
        return z
    # eval with autograd: plain tensor expression (not a hot path; inference uses conv_fused_eval)
    scale = gamma * torch.rsqrt(rv_p + bn.eps)
    pre = y * scale + (beta - rm_p * scale)
    if act == ops.ACT_RELU:
        pre = torch.relu(pre)
    elif act == ops.ACT_PRELU:
        pre = torch.where(pre > 0, pre, pre * slope)
    return to_operand(pre) if round_out else pre


# ----------------------------------------------------------------------------------------------- pad + concat
class PadCat(torch.autograd.Function):
    """Reflect-padded concatenation of NHWC maps (nearest-resized to the first map's size when they differ)."""

    @staticmethod
    def forward(ctx, pad, H, W, *srcs):
        N = srcs[0].shape[0]
        Ct = sum(s.shape[3] for s in srcs)
        buf = torch.empty(N, H + 2 * pad, W + 2 * pad, Ct, device=srcs[0].device, dtype=torch.float32)
        coff = 0
        meta = []
        for s in srcs:
            _, Hs, Ws, Cs = s.shape
            ops.copy_view(s, ops.view8(Hs, Ws, ld=Cs), buf, ops.view8(H, W, H + 2 * pad, W + 2 * pad, pad, pad, Ct, coff), N, Cs)
            meta.append((Hs, Ws, Cs, coff))
            coff += Cs
        if pad:
            ops.reflect_fill(buf, H, W, pad)
        ctx.meta, ctx.pad, ctx.H, ctx.W, ctx.N, ctx.Ct = meta, pad, H, W, N, Ct
        return buf

    @staticmethod
    def backward(ctx, gbuf):
        pad, H, W, N, Ct = ctx.pad, ctx.H, ctx.W, ctx.N, ctx.Ct
        g = gbuf.contiguous()
        folded = pad == 0
        grads = []
        for i, (Hs, Ws, Cs, coff) in enumerate(ctx.meta):
            if not ctx.needs_input_grad[3 + i]:
                grads.append(None)
                continue
            dv = ops.view8(H, W, H + 2 * pad, W + 2 * pad, pad, pad, Ct, coff)
            if (Hs, Ws) == (H, W):
                gs = torch.empty(N, Hs, Ws, Cs, device=g.device, dtype=torch.float32)
                if folded:
                    ops.copy_view(g, dv, gs, ops.view8(Hs, Ws, ld=Cs), N, Cs)
                else:                                              # fold + un-pad + slice in one pass, the padded map stays as it is
                    ops.copy_view_fold(g, dv, gs, ops.view8(Hs, Ws, ld=Cs), N, Cs)
            else:
                if not folded:                                     # (a resized source: fold a copy in place first, once)
                    g = g.clone()
                    ops.reflect_fold(g, H, W, pad)
                    folded = True
                gs = torch.zeros(N, Hs, Ws, Cs, device=g.device, dtype=torch.float32)
                ops.copy_view_backward(g, dv, gs, ops.view8(Hs, Ws, ld=Cs), N, Cs)
            grads.append(gs)
        return (None, None, None, *grads)


class PadCatH(torch.autograd.Function):
    """PadCat over half-map handles: the copies move 16-byte pixel groups, so they run on the fp32-pair view of the same bytes
    (channel counts halved); gradients are ordinary fp32 maps."""

    @staticmethod
    def forward(ctx, pad, H, W, *srcs):
        N = srcs[0].shape[0]
        Ct = sum(s.shape[3] for s in srcs)
        buf = ops.new_half((N, H + 2 * pad, W + 2 * pad, Ct), srcs[0].device)
        bf = ops.fv(buf)
        coff = 0
        meta = []
        for s in srcs:
            _, Hs, Ws, Cs = s.shape
            ops.copy_view(ops.fv(s), ops.view8(Hs, Ws, ld=Cs // 2), bf, ops.view8(H, W, H + 2 * pad, W + 2 * pad, pad, pad, Ct // 2, coff // 2),
                          N, Cs // 2)
            meta.append((Hs, Ws, Cs, coff))
            coff += Cs
        if pad:
            ops.reflect_fill(bf, H, W, pad)
        ctx.meta, ctx.pad, ctx.H, ctx.W, ctx.N, ctx.Ct = meta, pad, H, W, N, Ct
        return buf

    backward = staticmethod(PadCat.backward)


# ----------------------------------------------------------------------------------------------- layout changes
class ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, channels):
        ctx.cp = x.shape[3]
        return ops.nhwc_to_nchw(x, channels)

    @staticmethod
    def backward(ctx, g):
        return ops.nchw_to_nhwc(g.contiguous(), ctx.cp), None


class FeatToSeq(torch.autograd.Function):
    """Encoder outputs (B,F,T,C_i) NHWC -> LSTM input (V, B, sum C_i*F), feature = c*F + f (per source, concatenated)."""

    @staticmethod
    def forward(ctx, V, real_channels, *srcs):
        B, F, T, _ = srcs[0].shape
        ld = sum(c * F for c in real_channels)
        out = torch.empty(V, B, ld, device=srcs[0].device, dtype=torch.float32)
        coff = 0
        ctx.meta = []
        for s, c in zip(srcs, real_channels):
            # only the first c channels are real: view the padded NHWC map as C = Cp and let the kernel skip the rest
            sc = s if s.shape[3] == c else s[..., :c].contiguous()
            ops.feat_to_seq(sc, out, V, coff)
            ctx.meta.append((tuple(s.shape), c, coff))
            coff += c * F
        ctx.V = V
        return out

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous()
        grads = []
        for shape, c, coff in ctx.meta:
            B, F, T, Cp = shape
            g = ops.feat_to_seq_backward(gout, (B, F, T, c), ctx.V, coff)
            if Cp != c:
                g = torch.nn.functional.pad(g, (0, Cp - c))
            grads.append(g)
        return (None, None, *grads)


# ----------------------------------------------------------------------------------------------- Linear / LSTM
def _direct(p):
    """Inside the agents' backward (async_wgrad) a parameter gradient is added straight into p.grad (a view of the flat gradient
    buffer) by this library's kernels and autograd gets None."""
    return _ASYNC_WGRAD and _DIRECT_GRADS and p.grad is not None and p.grad.is_contiguous()


def _grad_out(p, g):
    if _direct(p):
        ops.axpy_(p.grad, g.contiguous() if not g.is_contiguous() else g)
        return None
    return g


_GRAD_3PASS = os.environ.get("SOS_GRAD_3PASS", "0") == "1"      # A/B switch: fp32-grade (3-pass) gradient GEMMs of the LSTM / MLP head


def _bw_passes():
    """Passes of the split-TF32 GEMMs in BACKWARD: the forward projections feed the sigmoid mask and stay fp32-grade (three passes);
    their data and weight gradients run as one TF32 pass in the default (half) mode -- the 11-bit-significand contract every
    convolution gradient of this mode already has (cuDNN's own default for the reference's LSTM) -- and as three in precise mode."""
    return 3 if (precise() or not half_mode() or _GRAD_3PASS) else 1


class LinearAct(torch.autograd.Function):
    """act(x @ W^T + b) over the rows of x -- nn.Linear + ReLU / Sigmoid of the heads (M1/networks.py:96-98, M2/networks.py:65-70)
    at fp32-grade accuracy on the tensor cores (ops.gemm3: one 3-tap TF32 tap-GEMM launch over split operands; bias and activation
    in its epilogue).  Backward: dpre = dy act'(y) with the bias gradient in the same pass, then two more gemm3 launches."""

    @staticmethod
    def forward(ctx, x, W, b, act):
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        y = ops.gemm3(ops.split_act(x2), ops.split_weight(W.detach()), W.shape[0], b.detach(), act, tag="linear")
        if not y.is_contiguous():
            y = y.contiguous()
        ctx.act = act
        ctx.save_for_backward(x2, W, b, y)
        return y.view(*x.shape[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, W, b, y = ctx.saved_tensors
        n_out, K = W.shape
        dy2 = dy.reshape(-1, n_out)
        dy2 = dy2 if dy2.is_contiguous() else dy2.contiguous()
        direct_b = _direct(b)
        db = b.grad if direct_b else torch.zeros(n_out, device=dy.device, dtype=torch.float32)
        dpre = ops.bias_act_backward(dy2, y, ctx.act, db)
        dx = dW = None
        np_ = _bw_passes()
        if ctx.needs_input_grad[0]:
            dx = ops.gemm3(ops.split_act(dpre, passes=np_), ops.split_weight(W.detach(), transpose=True, passes=np_), K, tag="linear_dgrad")
            dx = (dx if dx.is_contiguous() else dx.contiguous()).view(*dy.shape[:-1], K)
        if ctx.needs_input_grad[1]:
            dW = ops.gemm3(ops.split_act(dpre, transpose=True, passes=np_), ops.split_weight(x2, transpose=True, passes=np_), K, tag="linear_wgrad")
            dW = _grad_out(W, dW)
        return dx, dW, (None if direct_b else db), None


class SeqToMap(torch.autograd.Function):
    """(T, B, C) sequence rows -> (B, C, T): `h.permute(0, 2, 1).view(B, 2, 256, T)` of M2/networks.py:92-93 on (T, B)-ordered rows."""

    @staticmethod
    def forward(ctx, h):
        ctx.shape = h.shape
        return ops.seq_to_map(h.contiguous(), h.shape[2])

    @staticmethod
    def backward(ctx, g):
        T, B, Cn = ctx.shape
        return ops.map_to_seq(g.contiguous(), T, B, Cn)


class BiLSTMFn(torch.autograd.Function):
    """Single-layer bidirectional LSTM (nn.LSTM(bidirectional=True), M1/networks.py:95, M2/networks.py:64).  The input projection and
    every weight gradient are fp32-grade tensor-core GEMMs of this library (ops.gemm3); the recurrence runs in the sos_lstm_* kernels."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        T, B, I = x.shape
        H = w_hh.shape[1]
        x2 = x.reshape(T * B, I)
        a2 = ops.split_act(x2)
        gx = torch.empty(T * B, 8 * H, device=x.device, dtype=torch.float32)
        for d, (w, bi, bh) in enumerate(((w_ih, b_ih, b_hh), (w_ih_r, b_ih_r, b_hh_r))):
            bias = ops.axpy_(torch.empty(4 * H, device=x.device, dtype=torch.float32), bh.detach(), 1.0, base=bi.detach())
            ops.gemm3(a2, ops.split_weight(w.detach()), 4 * H, bias, 0, tag="lstm_proj", out=gx, col=d * 4 * H)
        whh = torch.empty(2, 4 * H, H, device=x.device, dtype=torch.float32)
        ops.axpy_(whh[0], w_hh.detach(), 0.0, base=w_hh.detach())               # (copies: whh = stack([w_hh, w_hh_r]))
        ops.axpy_(whh[1], w_hh_r.detach(), 0.0, base=w_hh_r.detach())
        out, gates, cell = ops.lstm_forward(gx.view(T, B, 2, 4 * H), whh)
        ctx.save_for_backward(x2, w_ih, w_ih_r, whh, out, gates, cell, w_hh, w_hh_r, b_ih, b_hh, b_ih_r, b_hh_r)
        ctx.dims = (T, B, I, H)
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, w_ih, w_ih_r, whh, out, gates, cell, w_hh, w_hh_r, b_ih, b_hh, b_ih_r, b_hh_r = ctx.saved_tensors
        T, B, I, H = ctx.dims
        M = T * B
        dgx = ops.lstm_backward(dout.contiguous(), whh, out, gates, cell)   # (T,B,2,4H)
        flat = dgx.view(M, 8 * H)
        dx = None
        np_ = _bw_passes()
        if ctx.needs_input_grad[0]:
            # dx = dg_f @ w_ih + dg_r @ w_ih_r: ONE GEMM over K = 8H against [w_ih; w_ih_r]^T
            dx = ops.gemm3(ops.split_act(flat, passes=np_), ops.split_weight_cat_t([w_ih.detach(), w_ih_r.detach()], passes=np_), I,
                           tag="lstm_dgrad").view(T, B, I)
        xT = ops.split_weight(x2, transpose=True, passes=np_)               # (I, passes * MP): shared by both directions
        out2 = out.view(M, 2 * H)
        grads = []
        for d, (wi, wh, bi, bh) in enumerate(((w_ih, w_hh, b_ih, b_hh), (w_ih_r, w_hh_r, b_ih_r, b_hh_r))):
            dg = flat[:, d * 4 * H:(d + 1) * 4 * H]                         # (M, 4H), row stride 8H
            dgT = ops.split_act(dg, transpose=True, passes=np_)             # (2 | 1, 4H, MP)
            dW = _grad_out(wi, ops.gemm3(dgT, xT, I, tag="lstm_wgrad"))
            # h_{t-1} of the forward direction is out[t-1, :, :H] (rows shifted by -B); of the reverse direction out[t+1, :, H:]
            hprevT = ops.split_weight(out2[:, d * H:(d + 1) * H], transpose=True, k_shift=(-B if d == 0 else B), passes=np_)
            dWhh = ops.gemm3(dgT, hprevT, H, tag="lstm_wgrad")
            dWhh = _grad_out(wh, dWhh if dWhh.is_contiguous() else dWhh.contiguous())
            if _direct(bi) and _direct(bh):
                ops.bias_act_backward(dg, dg, 0, bi.grad, want_dpre=False)
                ops.bias_act_backward(dg, dg, 0, bh.grad, want_dpre=False)
                db_i = db_h = None
            else:
                db_i = torch.zeros(4 * H, device=dg.device, dtype=torch.float32)
                ops.bias_act_backward(dg, dg, 0, db_i, want_dpre=False)
                db_h = db_i
            grads += [dW, dWhh, db_i, db_h]
        return (dx, *grads)


# ----------------------------------------------------------------------------------------------- cRM + losses
class ICRM(torch.autograd.Function):
    """batch_fast_icRM_sigmoid (M2/transform.py:156-169); gradient flows to the mask only."""

    @staticmethod
    def forward(ctx, Y, crm, a, b):
        Y, crm = Y.contiguous(), crm.contiguous()
        ctx.a = a
        ctx.save_for_backward(Y, crm)
        return ops.icrm_forward(Y, crm, a, b)

    @staticmethod
    def backward(ctx, grec):
        Y, crm = ctx.saved_tensors
        return None, ops.icrm_backward(Y, crm, grec.contiguous(), ctx.a), None, None


class MSELoss(torch.autograd.Function):
    """nn.MSELoss() (mean); loss and gradient come out of one pass."""

    @staticmethod
    def forward(ctx, pred, target):
        loss, grad = ops.mse_fwd_bwd(pred.contiguous(), target.contiguous(), ctx.needs_input_grad[0])
        ctx.save_for_backward(grad if grad is not None else torch.empty(0))
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


class BCEWithLogitsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        loss, grad = ops.bce_fwd_bwd(logits.contiguous(), labels.contiguous(), ctx.needs_input_grad[0])
        ctx.save_for_backward(grad if grad is not None else torch.empty(0))
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None
