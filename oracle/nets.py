"""Functional torch-CPU fp32 restatement of the reference networks
(TEST INFRASTRUCTURE).

The networks are interpreted from a layer table over a plain ``state_dict``
(reference key names), so the same weights drive the reference module, this
oracle and the CUDA path.

  SID   AudioVisualNet            M1/networks.py:80-155  (Conv2dBlock :28-51)
  Joint JointModel                M2/networks.py:208-217
        InpaintNet                M2/networks.py:152-205 (Down/UpConvBlock :97-149)
        ContextAggNet             M2/networks.py:54-94   (ConvBlock :28-51)
  hyper-parameters                M2/common.py:80-81, M1/networks.py:91-93
"""
import torch
import torch.nn.functional as F

KS = [(1, 7), (7, 1)] + [(5, 5)] * 12                      # M2/common.py:80
DL = [(1, 1), (1, 1), (1, 1), (2, 1), (4, 1), (8, 1), (16, 1), (32, 1),
      (1, 1), (2, 2), (4, 4), (8, 8), (16, 16), (32, 32)]  # M2/common.py:81
SID_KS, SID_DL = KS[:11], DL[:11]                          # M1/networks.py:91-92
BN_EPS, BN_MOM = 1e-5, 0.1

# ---------------------------------------------------------------------------
# Arithmetic contract of the convolutions.
#   TF32 = False : plain fp32 (what the golden vectors from the reference hold)
#   TF32 = True  : "TF32 contract" -- every convolution multiplies operands rounded
#                  (to nearest, ties away: PTX cvt.rna.tf32.f32) to a 10-bit mantissa and
#                  accumulates in fp32, forward AND backward (data and weight gradients
#                  see the rounded output gradient).  This is what the tcgen05
#                  kind::tf32 kernels compute, and what PyTorch's own cuDNN path does
#                  for the reference's nn.Conv2d on Ampere+ GPUs (allow_tf32 = True).
#   half_contract: the same contract with round-to-nearest-EVEN (what cvt.rn.f16.f32 does):
#                  the tcgen05 kind::f16 kernels multiply IEEE half operands, which carry the
#                  same 11-bit significand as TF32; gradient operands travel scaled by a
#                  power of two, so the rounding is range-free here.
# ---------------------------------------------------------------------------
TF32 = False
RNE = False


class tf32_contract(object):
    def __enter__(self):
        global TF32
        self.old, TF32 = TF32, True

    def __exit__(self, *a):
        global TF32
        TF32 = self.old


class half_contract(object):
    def __enter__(self):
        global TF32, RNE
        self.old, TF32, RNE = (TF32, RNE), True, True

    def __exit__(self, *a):
        global TF32, RNE
        TF32, RNE = self.old


def round_tf32(t):
    """fp32 -> nearest value with a 10-bit mantissa, kept in fp32: ties away from zero (TF32, cvt.rna), or ties to even
    under half_contract (cvt.rn.f16 inside half's normal range)."""
    i = t.detach().contiguous().view(torch.int32)
    if RNE:
        return ((i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _TF32Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, dilation, padding):
        xr, wr = round_tf32(x), round_tf32(w)
        ctx.save_for_backward(xr, wr)
        ctx.cfg = (stride, dilation, padding)
        return F.conv2d(xr, wr, None, stride, padding, dilation)

    @staticmethod
    def backward(ctx, dy):
        xr, wr = ctx.saved_tensors
        stride, dilation, padding = ctx.cfg
        dyr = round_tf32(dy)
        dx = torch.nn.grad.conv2d_input(xr.shape, wr, dyr, stride, padding, dilation)
        dw = torch.nn.grad.conv2d_weight(xr, wr.shape, dyr, stride, padding, dilation)
        return dx, dw, None, None, None


class _TF32ConvT(torch.autograd.Function):
    """ConvTranspose2d(k3, s2, p1, output_padding=1) under the same contract."""

    @staticmethod
    def forward(ctx, x, w):
        xr, wr = round_tf32(x), round_tf32(w)
        ctx.save_for_backward(xr, wr)
        return F.conv_transpose2d(xr, wr, None, 2, 1, 1)

    @staticmethod
    def backward(ctx, dy):
        xr, wr = ctx.saved_tensors
        dyr = round_tf32(dy)
        dx = F.conv2d(dyr, wr, None, 2, 1)                      # the transposed conv is the data gradient of this conv
        dw = torch.nn.grad.conv2d_weight(dyr, wr.shape, xr, 2, 1)
        return dx, dw


def _conv(x, w, stride=1, padding=0, dilation=1):
    if TF32:
        return _TF32Conv.apply(x, w, stride, dilation, padding)
    return F.conv2d(x, w, None, stride, padding, dilation)


def _convT(x, w):
    if TF32:
        return _TF32ConvT.apply(x, w)
    return F.conv_transpose2d(x, w, None, 2, 1, 1)


def _bn(sd, pfx, y, training, stats_out):
    rm, rv = sd[pfx + ".running_mean"], sd[pfx + ".running_var"]
    if training:
        # batch statistics; running buffers are updated on copies handed back
        rm2, rv2 = rm.clone(), rv.clone()
        out = F.batch_norm(y, rm2, rv2, sd[pfx + ".weight"], sd[pfx + ".bias"], True, BN_MOM, BN_EPS)
        if stats_out is not None:
            stats_out[pfx + ".running_mean"] = rm2
            stats_out[pfx + ".running_var"] = rv2
        return out
    return F.batch_norm(y, rm, rv, sd[pfx + ".weight"], sd[pfx + ".bias"], False, BN_MOM, BN_EPS)


def _zero_block(sd, pfx, x, k, d, training, stats_out):
    """Conv(bias=False, zero 'same' pad) + BN + ReLU  (ConvBlock / Conv2dBlock)."""
    pad = ((k[0] - 1) // 2 * d[0], (k[1] - 1) // 2 * d[1])
    y = _conv(x, sd[pfx + ".block.0.weight"], 1, pad, d)
    return F.relu(_bn(sd, pfx + ".block.1", y, training, stats_out))


def _encoder(sd, pfx, x, ks, dl, training, stats_out):
    for i, (k, d) in enumerate(zip(ks, dl)):
        x = _zero_block(sd, f"{pfx}.{i}", x, k, d, training, stats_out)
    return _zero_block(sd, f"{pfx}.{len(ks)}", x, (1, 1), (1, 1), training, stats_out)


def _lstm(sd, pfx, x, hidden):
    """Bidirectional single-layer LSTM, PyTorch gate order i,f,g,o; x (T,B,I)."""
    outs = []
    for sfx, rev in (("", False), ("_reverse", True)):
        w_ih, w_hh = sd[f"{pfx}.weight_ih_l0{sfx}"], sd[f"{pfx}.weight_hh_l0{sfx}"]
        b = sd[f"{pfx}.bias_ih_l0{sfx}"] + sd[f"{pfx}.bias_hh_l0{sfx}"]
        gx = x @ w_ih.t() + b
        h = x.new_zeros(x.shape[1], hidden)
        c = x.new_zeros(x.shape[1], hidden)
        hs = [None] * x.shape[0]
        order = range(x.shape[0] - 1, -1, -1) if rev else range(x.shape[0])
        for t in order:
            g = gx[t] + h @ w_hh.t()
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            hs[t] = h
        outs.append(torch.stack(hs))
    return torch.cat(outs, dim=2)


def sid_forward(sd, s, v_num_frames=60, training=False, stats_out=None):
    """AudioVisualNet.forward  M1/networks.py:130-155.  s (B,2,256,T) -> (B,v)."""
    f = _encoder(sd, "encoder_audio", s, SID_KS, SID_DL, training, stats_out)
    f = f.reshape(f.size(0), -1, f.size(3))
    f = F.interpolate(f, size=v_num_frames)              # nearest: src = floor(i*T/v)
    m = _lstm(sd, "lstm", f.permute(2, 0, 1), 100).permute(1, 0, 2)
    m = F.relu(m @ sd["fc1.0.weight"].t() + sd["fc1.0.bias"])
    m = m @ sd["fc1.2.weight"].t() + sd["fc1.2.bias"]
    return m.squeeze(2)


def _down(sd, pfx, x, k, stride, d, training, stats_out, norm=True):
    """ReflectionPad + Conv + BN + PReLU  (DownConvBlock M2/networks.py:97-117)."""
    p = (k - 1) // 2 * d
    x = F.pad(x, (p, p, p, p), mode="reflect")
    if not norm:
        return _conv(x, sd[pfx + ".block.1.weight"], stride, 0, d) + sd[pfx + ".block.1.bias"].view(1, -1, 1, 1)
    y = _conv(x, sd[pfx + ".block.1.weight"], stride, 0, d)
    return F.prelu(_bn(sd, pfx + ".block.2", y, training, stats_out), sd[pfx + ".block.3.weight"])


def _up(sd, pfx, x, training, stats_out):
    """ConvTranspose2d(k3,s2,p1,output_padding=1) + BN + PReLU (UpConvBlock :120-149;
    `dilation` lands in the output_padding slot, :130)."""
    y = _convT(x, sd[pfx + ".block.0.weight"])
    return F.prelu(_bn(sd, pfx + ".block.1", y, training, stats_out), sd[pfx + ".block.2.weight"])


def inpaint_forward(sd, pfx, x, y, training=False, stats_out=None):
    """InpaintNet.forward(x, y)  M2/networks.py:192-205."""
    a = (training, stats_out)
    d1 = _down(sd, pfx + "down1.0", x, 5, 1, 1, *a)
    d2 = _down(sd, pfx + "down2.1", _down(sd, pfx + "down2.0", d1, 5, 2, 1, *a), 5, 1, 1, *a)
    d3 = _down(sd, pfx + "down3.0", y, 5, 1, 1, *a)
    d4 = _down(sd, pfx + "down4.1", _down(sd, pfx + "down4.0", d3, 5, 2, 1, *a), 5, 1, 1, *a)
    o = torch.cat([d2, d4], dim=1)
    o = _down(sd, pfx + "mid.0", o, 3, 2, 1, *a)
    for i, d in zip(range(1, 8), (1, 2, 4, 8, 16, 1, 1)):
        o = _down(sd, f"{pfx}mid.{i}", o, 3, 1, d, *a)
    o = _up(sd, pfx + "mid.8", o, *a)
    if o.shape != d4.shape:
        o = F.interpolate(o, d4.shape[-2:])
    o = _down(sd, pfx + "up1.0", torch.cat([o, d4], dim=1), 3, 1, 1, *a)
    o = _up(sd, pfx + "up1.1", o, *a)
    if o.shape != d3.shape:
        o = F.interpolate(o, d3.shape[-2:])
    o = _down(sd, pfx + "up2.0", torch.cat([o, d3], dim=1), 3, 1, 1, *a)
    return _down(sd, pfx + "up2.1", o, 3, 1, 1, *a, norm=False)


def context_forward(sd, pfx, x, n, training=False, stats_out=None):
    """ContextAggNet.forward(x, n)  M2/networks.py:82-94."""
    fx = _encoder(sd, pfx + "encoder_x", x, KS, DL, training, stats_out)
    fx = fx.reshape(fx.size(0), -1, fx.size(3)).permute(2, 0, 1)
    fn = _encoder(sd, pfx + "encoder_n", n, KS, DL, training, stats_out)
    fn = fn.reshape(fn.size(0), -1, fn.size(3)).permute(2, 0, 1)
    h = _lstm(sd, pfx + "lstm", torch.cat([fx, fn], dim=2), 200).permute(1, 0, 2)
    h = F.relu(h @ sd[pfx + "fc.0.weight"].t() + sd[pfx + "fc.0.bias"])
    h = F.relu(h @ sd[pfx + "fc.2.weight"].t() + sd[pfx + "fc.2.bias"])
    h = torch.sigmoid(h @ sd[pfx + "fc.4.weight"].t() + sd[pfx + "fc.4.bias"])
    return h.permute(0, 2, 1).reshape(h.size(0), 2, -1, h.size(1))


def joint_forward(sd, x, n, training=False, stats_out=None):
    """JointModel.forward(x, n)  M2/networks.py:214-217 (note stage1(n, x))."""
    n_pred = inpaint_forward(sd, "stage1.", n, x, training, stats_out)
    mask = context_forward(sd, "stage2.", x, n_pred, training, stats_out)
    return n_pred, mask


# ---------------------------------------------------------------------------
# state_dict shapes (reference key names) and a deterministic weight filler
# ---------------------------------------------------------------------------
def _enc_shapes(pfx, ks, nf, outf, shapes):
    cin = 2
    for i, k in enumerate(list(ks) + [(1, 1)]):
        cout = outf if i == len(ks) else nf
        shapes[f"{pfx}.{i}.block.0.weight"] = (cout, cin, k[0], k[1])
        _bn_shapes(f"{pfx}.{i}.block.1", cout, shapes)
        cin = cout


def _bn_shapes(pfx, c, shapes):
    shapes[pfx + ".weight"] = (c,)
    shapes[pfx + ".bias"] = (c,)
    shapes[pfx + ".running_mean"] = (c,)
    shapes[pfx + ".running_var"] = (c,)
    shapes[pfx + ".num_batches_tracked"] = ()


def _lstm_shapes(pfx, inp, hid, shapes):
    for sfx in ("", "_reverse"):
        shapes[f"{pfx}.weight_ih_l0{sfx}"] = (4 * hid, inp)
        shapes[f"{pfx}.weight_hh_l0{sfx}"] = (4 * hid, hid)
        shapes[f"{pfx}.bias_ih_l0{sfx}"] = (4 * hid,)
        shapes[f"{pfx}.bias_hh_l0{sfx}"] = (4 * hid,)


def sid_shapes():
    s = {}
    _enc_shapes("encoder_audio", SID_KS, 48, 8, s)
    _lstm_shapes("lstm", 2048, 100, s)
    s["fc1.0.weight"], s["fc1.0.bias"] = (100, 200), (100,)
    s["fc1.2.weight"], s["fc1.2.bias"] = (1, 100), (1,)
    return s


def joint_shapes():
    s = {}

    def down(pfx, cin, cout, k, norm=True):
        s[pfx + ".block.1.weight"] = (cout, cin, k, k)
        if norm:
            _bn_shapes(pfx + ".block.2", cout, s)
            s[pfx + ".block.3.weight"] = (1,)
        else:
            s[pfx + ".block.1.bias"] = (cout,)

    def up(pfx, cin, cout):
        s[pfx + ".block.0.weight"] = (cin, cout, 3, 3)
        _bn_shapes(pfx + ".block.1", cout, s)
        s[pfx + ".block.2.weight"] = (1,)

    p = "stage1."
    down(p + "down1.0", 2, 64, 5); down(p + "down2.0", 64, 128, 5); down(p + "down2.1", 128, 128, 5)
    down(p + "down3.0", 2, 64, 5); down(p + "down4.0", 64, 128, 5); down(p + "down4.1", 128, 128, 5)
    for i in range(8):
        down(f"{p}mid.{i}", 256, 256, 3)
    up(p + "mid.8", 256, 128)
    down(p + "up1.0", 256, 128, 3); up(p + "up1.1", 128, 64)
    down(p + "up2.0", 128, 64, 3); down(p + "up2.1", 64, 2, 3, norm=False)
    _enc_shapes("stage2.encoder_x", KS, 96, 8, s)
    _enc_shapes("stage2.encoder_n", KS, 48, 4, s)
    _lstm_shapes("stage2.lstm", 3072, 200, s)
    s["stage2.fc.0.weight"], s["stage2.fc.0.bias"] = (600, 400), (600,)
    s["stage2.fc.2.weight"], s["stage2.fc.2.bias"] = (600, 600), (600,)
    s["stage2.fc.4.weight"], s["stage2.fc.4.bias"] = (512, 600), (512,)
    return s


def synth_state_dict(shapes, seed, spread=False):
    """Deterministic weights keyed by name: fan-in scaled uniform for matrices,
    non-trivial BN affine / running stats, PReLU 0.25.  `spread` widens the
    final mask layer so the sigmoid output covers (0,1)."""
    import zlib
    sd = {}
    for name, shp in shapes.items():
        g = torch.Generator().manual_seed(seed * 1000003 + (zlib.crc32(name.encode()) & 0x7FFFFFFF))
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith("running_mean"):
            sd[name] = 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shp, generator=g)
        elif len(shp) == 1 and shp[0] == 1 and ".block." in name and not name.endswith("bias"):
            sd[name] = torch.full(shp, 0.25)                       # PReLU slope
        elif len(shp) == 1 and name.endswith(".weight") and ".block." in name:
            sd[name] = 0.75 + 0.5 * torch.rand(shp, generator=g)   # BN gamma
        elif len(shp) == 1:
            sd[name] = 0.1 * (torch.rand(shp, generator=g) - 0.5)  # biases, BN beta
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            bound = (3.0 / fan_in) ** 0.5
            sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    if spread and "stage2.fc.4.weight" in sd:
        sd["stage2.fc.4.weight"] = sd["stage2.fc.4.weight"] * 20
    return sd
