"""The two networks of the hot path with the reference's class names, constructor order, forward signatures and
state_dict keys, running on the sm_100a kernels (NHWC fp32 activations, tcgen05 tap GEMMs).

  AudioVisualNet  (SID)            M1/networks.py:80-155      get_network()        M1/networks.py:8-9
  InpaintNet / ContextAggNet /
  JointModel                       M2/networks.py:54-217      get_network(config)  M2/networks.py:8-9

The sub-modules are ordinary nn.Conv2d / nn.BatchNorm2d / nn.PReLU / nn.LSTM / nn.Linear instances created in the
reference's order (so `torch.manual_seed(s); get_network()` gives bit-identical initial weights and the reference's
checkpoints load with `load_state_dict`), but they are used as parameter holders only: forward() never calls them
except for the small nn.Linear heads (plain library GEMMs).
"""
import torch
import torch.nn as nn

from . import layers as L
from . import ops

KERNEL_SIZES = [(1, 7), (7, 1)] + [(5, 5)] * 12                               # M2/common.py:80
DILATIONS = [(1, 1), (1, 1), (1, 1), (2, 1), (4, 1), (8, 1), (16, 1), (32, 1),
             (1, 1), (2, 2), (4, 4), (8, 8), (16, 16), (32, 32)]              # M2/common.py:81


def _fast_eval(module, *tensors):
    """Inference path (BN folded into the GEMM epilogue): eval mode with autograd off."""
    return (not module.training) and not torch.is_grad_enabled() and not L.precise()


class _Block(nn.Module):
    """Common driver: conv (tap GEMM) -> BatchNorm -> activation over NHWC maps."""
    round_out = True          # the block's output feeds another tensor-core GEMM: round it to TF32 where it is produced

    def _run(self, x, conv, bn, act, slope_mod, geom):
        slope = slope_mod.weight if slope_mod is not None else None
        half = ops.is_half_handle(x)                                     # half-map handle (default mode) or fp32 / TF32 map
        if bn is None:                                                   # last InpaintNet conv: bias, no norm, no act
            y = (L.TapConvH if half else L.TapConv).apply(x, conv.weight, geom)
            Cp, Cn = y.shape[3], conv.out_channels
            b = conv.bias if Cp == Cn else torch.nn.functional.pad(conv.bias, (0, Cp - Cn))
            return y + b
        if _fast_eval(self, x):
            gamma, beta, rm, rv = bn.weight, bn.bias, bn.running_mean, bn.running_var
            Cp = (bn.num_features + 7) // 8 * 8
            if Cp != bn.num_features:
                pad = (0, Cp - bn.num_features)
                gamma, beta, rm = (torch.nn.functional.pad(t, pad) for t in (gamma, beta, rm))
                rv = torch.nn.functional.pad(rv, pad, value=1.0)
            scale, shift = ops.bn_eval_coeffs(gamma.detach().contiguous(), beta.detach().contiguous(), rm.contiguous(), rv.contiguous(), bn.eps)
            return L.conv_fused_eval(x, conv.weight.detach(), geom, scale, shift, act, slope.detach() if slope is not None else None,
                                     self.round_out)
        if self.training and half:                                       # conv + batch-stat BN + activation as one autograd node
            z = L.ConvBNActH.apply(x, conv.weight, bn.weight, bn.bias, slope, bn.running_mean, bn.running_var, bn.eps, bn.momentum,
                                   act, geom, not self.round_out)
            L.note_bn_step(bn)
            return z
        if self.training:                                                # batch statistics come out of the conv epilogue
            y, partial = L.TapConv.apply(x, conv.weight, geom, True)
            return L.bn_act(y, bn, act, slope, True, self.round_out, conv_partial=partial)
        y = (L.TapConvH if half else L.TapConv).apply(x, conv.weight, geom)
        return L.bn_act(y, bn, act, slope, False, self.round_out)


class ConvBlock(_Block):
    """Conv2d(bias=False, zero 'same' padding, dilation) + BatchNorm2d + ReLU  (M2/networks.py:28-51, M1/networks.py:28-51)."""

    def __init__(self, in_channels, out_channels, kernel_size, dilation):
        super().__init__()
        pad = ((kernel_size[0] - 1) // 2 * dilation[0], (kernel_size[1] - 1) // 2 * dilation[1])
        self.block = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size, 1, pad, dilation, bias=False),
                                   nn.BatchNorm2d(out_channels), nn.ReLU())
        self.geom = L.ConvGeom("zero", kernel_size[0], kernel_size[1], dilation[0], dilation[1], 1)

    def forward(self, x):
        return self._run(x, self.block[0], self.block[1], ops.ACT_RELU, None, self.geom)


Conv2dBlock = ConvBlock


class DownConvBlock(_Block):
    """ReflectionPad2d + Conv2d(stride, dilation) + BatchNorm2d + PReLU  (M2/networks.py:97-117).
    forward takes the list of NHWC maps that the reference concatenates (and resizes) before the block."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, dilation=1, norm_fn='bn', act='prelu'):
        super().__init__()
        self.pad = (kernel_size - 1) // 2 * dilation
        block = [nn.ReflectionPad2d(self.pad),
                 nn.Conv2d(in_channels, out_channels, kernel_size, stride, 0, dilation, bias=norm_fn is None)]
        if norm_fn == 'bn':
            block.append(nn.BatchNorm2d(out_channels))
        if act == 'prelu':
            block.append(nn.PReLU())
        self.block = nn.Sequential(*block)
        self.has_norm = norm_fn == 'bn'
        self.geom = L.ConvGeom("valid", kernel_size, kernel_size, dilation, dilation, stride, round_dy=norm_fn != 'bn')

    def forward(self, *srcs, size=None):
        H, W = size if size is not None else (srcs[0].shape[1], srcs[0].shape[2])
        if self.pad >= min(H, W):
            raise RuntimeError(f"ReflectionPad2d({self.pad}) needs an input larger than the padding, got {H}x{W}")
        xp = (L.PadCatH if ops.is_half_handle(srcs[0]) else L.PadCat).apply(self.pad, H, W, *srcs)
        if self.has_norm:
            return self._run(xp, self.block[1], self.block[2], ops.ACT_PRELU, self.block[3], self.geom)
        return self._run(xp, self.block[1], None, ops.ACT_NONE, None, self.geom)


class UpConvBlock(_Block):
    """ConvTranspose2d(k, stride 2, padding, output_padding=dilation) + BatchNorm2d + PReLU  (M2/networks.py:120-149;
    the reference passes `dilation` positionally into the output_padding slot, :130)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, dilation=1):
        super().__init__()
        pad = (kernel_size - 1) // 2 * dilation
        self.block = nn.Sequential(nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, pad, dilation, bias=False),
                                   nn.BatchNorm2d(out_channels), nn.PReLU())
        assert (kernel_size, stride, pad, dilation) == (3, 2, 1, 1), "only the reference's k3 s2 p1 op1 transposed conv is built"
        self.geom = L.ConvGeom("convT", 3, 3, 1, 1, 2)

    def forward(self, x):
        return self._run(x, self.block[0], self.block[1], ops.ACT_PRELU, self.block[2], self.geom)


def _to_nhwc(x):
    """(B, 2, 256, T) NCHW -> NHWC operand of the first convolutions: a half map with 16 channels (one K = 16 MMA step) in the
    default mode, else fp32 with channels zero-padded to 8 (TF32-rounded unless precise)."""
    if L.half_mode():
        return ops.nchw_to_nhwc_half(x.contiguous().float(), 16)
    y = ops.nchw_to_nhwc(x.contiguous().float(), 8)
    return y if L.precise() else ops.round_tf32_(y)


class _NHWCInput(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _to_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_to_nchw(g.contiguous(), 2)


def _nhwc_in(x):
    return _NHWCInput.apply(x) if (torch.is_grad_enabled() and x.requires_grad) else _to_nhwc(x)


class BiLSTM(nn.LSTM):
    """nn.LSTM(input, hidden, bidirectional=True) parameters; recurrence on the sos_lstm kernels.  x (T, B, I)."""

    def forward(self, x):
        return L.BiLSTMFn.apply(x.contiguous(), self.weight_ih_l0, self.weight_hh_l0, self.bias_ih_l0, self.bias_hh_l0,
                                self.weight_ih_l0_reverse, self.weight_hh_l0_reverse, self.bias_ih_l0_reverse,
                                self.bias_hh_l0_reverse)


def _make_enc(kernel_sizes, dilations, nf, outf):
    enc = [ConvBlock(2 if i == 0 else nf, nf, kernel_sizes[i], dilations[i]) for i in range(len(kernel_sizes))]
    enc.append(ConvBlock(nf, outf, (1, 1), (1, 1)))
    enc[-1].round_out = False                                             # feeds the LSTM input projection (fp32 GEMM)
    return nn.Sequential(*enc)


_CHAIN = True                 # A/B switch: run training-mode encoders as one EncoderChainH node (half dz between the blocks)


def _run_encoder(enc, x):
    """An encoder Sequential of ConvBlocks on an NHWC operand map.  Training in the default (half) mode: the whole chain is one
    autograd node (layers.EncoderChainH); every other mode runs the blocks one by one."""
    if not (_CHAIN and enc.training and torch.is_grad_enabled() and ops.is_half_handle(x) and L.half_mode()):
        return enc(x)
    params = []
    for blk in enc:
        conv, bn = blk.block[0], blk.block[1]
        params += [conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
    bn0 = enc[0].block[1]
    z = L.EncoderChainH.apply(x, tuple(blk.geom for blk in enc), bn0.eps, bn0.momentum, *params)
    for blk in enc:
        L.note_bn_step(blk.block[1])
    return z


class AudioVisualNet(nn.Module):
    """Silent-interval detector (M1/networks.py:80-155).  forward(s (B,2,256,T), v_num_frames) -> logits (B, v)."""

    def __init__(self, freq_bins=256, time_bins=178, nf=96):
        super().__init__()
        self.encoder_audio = _make_enc(KERNEL_SIZES[:11], DILATIONS[:11], 48, 8)
        self.lstm = BiLSTM(input_size=8 * freq_bins, hidden_size=100, bidirectional=True)
        self.fc1 = nn.Sequential(nn.Linear(200, 100), nn.ReLU(True), nn.Linear(100, 1))

    def forward(self, s, v_num_frames=60):
        f = _run_encoder(self.encoder_audio, _nhwc_in(s))                               # (B, 256, T, 8)
        seq = L.FeatToSeq.apply(int(v_num_frames), (8,), f)               # (v, B, 2048)
        m = self.lstm(seq)                                                # (v, B, 200); the head acts per row: keep (v, B) order
        m = L.LinearAct.apply(m, self.fc1[0].weight, self.fc1[0].bias, ops.ACT_RELU)
        m = L.LinearAct.apply(m, self.fc1[2].weight, self.fc1[2].bias, ops.ACT_NONE)
        L.flush_bn_steps()
        return m.squeeze(2).t()                                           # (B, v)


class ContextAggNet(nn.Module):
    """Mask predictor (M2/networks.py:54-94).  forward(x, n) with NCHW or already-NHWC inputs -> mask (B,2,256,T)."""

    def __init__(self, kernel_sizes, dilations, freq_bins=256, nf=96):
        super().__init__()
        self.encoder_x = _make_enc(kernel_sizes, dilations, nf, 8)
        self.encoder_n = _make_enc(kernel_sizes, dilations, nf // 2, 4)
        self.lstm = BiLSTM(input_size=8 * freq_bins + 4 * freq_bins, hidden_size=200, bidirectional=True)
        self.fc = nn.Sequential(nn.Linear(400, 600), nn.ReLU(True), nn.Linear(600, 600), nn.ReLU(True),
                                nn.Linear(600, freq_bins * 2), nn.Sigmoid())

    def forward(self, x, n):
        return self.forward_nhwc(_nhwc_in(x), _nhwc_in(n))

    def forward_nhwc(self, x, n):
        """x, n: NHWC (B,256,T,8) maps whose first two channels are real/imag."""
        fx = _run_encoder(self.encoder_x, x)
        fn = _run_encoder(self.encoder_n, n)
        T = fx.shape[2]
        seq = L.FeatToSeq.apply(T, (8, 4), fx, fn)                        # (T, B, 3072)
        h = self.lstm(seq)                                                # (T, B, 400); the head acts per row: keep (T, B) order
        h = L.LinearAct.apply(h, self.fc[0].weight, self.fc[0].bias, ops.ACT_RELU)
        h = L.LinearAct.apply(h, self.fc[2].weight, self.fc[2].bias, ops.ACT_RELU)
        h = L.LinearAct.apply(h, self.fc[4].weight, self.fc[4].bias, ops.ACT_SIGMOID)       # (T, B, 512)
        m = L.SeqToMap.apply(h)                                           # (B, 512, T) = h.permute(1, 2, 0)
        L.flush_bn_steps()
        return m.view(m.size(0), 2, -1, m.size(2))


class InpaintNet(nn.Module):
    """Noise estimator (M2/networks.py:152-205).  forward(x, y) -> NHWC (B,256,T,8) map whose first 2 channels are the
    predicted full-noise spectrogram (JointModel converts it to NCHW)."""

    def __init__(self):
        super().__init__()
        ch1, ch2, ch3 = 64, 128, 256
        self.down1 = nn.Sequential(DownConvBlock(2, ch1, 5, 1))
        self.down2 = nn.Sequential(DownConvBlock(ch1, ch2, 5, 2), DownConvBlock(ch2, ch2, 5, 1))
        self.down3 = nn.Sequential(DownConvBlock(2, ch1, 5, 1))
        self.down4 = nn.Sequential(DownConvBlock(ch1, ch2, 5, 2), DownConvBlock(ch2, ch2, 5, 1))
        self.mid = nn.Sequential(DownConvBlock(ch2 * 2, ch3, 3, 2), DownConvBlock(ch3, ch3, 3, 1),
                                 DownConvBlock(ch3, ch3, 3, 1, dilation=2), DownConvBlock(ch3, ch3, 3, 1, dilation=4),
                                 DownConvBlock(ch3, ch3, 3, 1, dilation=8), DownConvBlock(ch3, ch3, 3, 1, dilation=16),
                                 DownConvBlock(ch3, ch3, 3, 1), DownConvBlock(ch3, ch3, 3, 1), UpConvBlock(ch3, ch2, 3, 2))
        self.up1 = nn.Sequential(DownConvBlock(ch2 * 2, ch2, 3, 1), UpConvBlock(ch2, ch1, 3, 2))
        self.up2 = nn.Sequential(DownConvBlock(ch1 * 2, ch1, 3, 1), DownConvBlock(ch1, 2, 3, 1, norm_fn=None, act=None))

    def forward_nhwc(self, x, y):
        d1 = self.down1(x)
        d2 = self.down2(d1)
        d3 = self.down3(y)
        d4 = self.down4(d3)
        o = self.mid[0](d2, d4)                                           # torch.cat([down2, down4], 1)
        for i in range(1, 9):
            o = self.mid[i](o)
        # reference: `if out.shape != skip.shape: out = F.interpolate(out, skip.size()[-2:])`, then cat([out, skip]);
        # PadCat nearest-resizes every source to `size` while it builds the padded buffer
        o = self.up1[0](o, d4, size=d4.shape[1:3])
        o = self.up1[1](o)
        o = self.up2[0](o, d3, size=d3.shape[1:3])
        o = self.up2[1](o)
        L.flush_bn_steps()
        return o

    def forward(self, x, y):
        return L.ToNCHW.apply(self.forward_nhwc(_nhwc_in(x), _nhwc_in(y)), 2)


class JointModel(nn.Module):
    """M2/networks.py:208-217: n_pred = stage1(n, x); out = stage2(x, n_pred); return n_pred, out."""

    def __init__(self, config=None):
        super().__init__()
        ks = getattr(config, "kernel_sizes", KERNEL_SIZES)
        dl = getattr(config, "dilations", DILATIONS)
        self.stage1 = InpaintNet()
        self.stage2 = ContextAggNet(ks, dl)

    def forward(self, x, n):
        xh, nh = _nhwc_in(x), _nhwc_in(n)
        n_pred_nhwc = self.stage1.forward_nhwc(nh, xh)
        out = self.stage2.forward_nhwc(xh, L.to_operand(n_pred_nhwc))
        return L.ToNCHW.apply(n_pred_nhwc, 2), out


def get_network(config=None):
    """get_network() -> SID (M1/networks.py:8-9); get_network(config) -> JointModel (M2/networks.py:8-9)."""
    return AudioVisualNet() if config is None else JointModel(config)
