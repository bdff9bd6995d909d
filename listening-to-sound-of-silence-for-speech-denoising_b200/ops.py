"""Thin tensor-level wrappers over the C ABI (include/sos_b200.h).

Every function takes CUDA fp32 tensors, enqueues kernels of libsos_b200.so on PyTorch's current stream and
returns tensors.  PyTorch is used for device memory and streams only; there is no fallback path.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ConvArgs, WgradArgs, check, lib

_I32 = C.c_int32


def _stream():
    # (torch.cuda.current_stream() builds a Python Stream object per call: ~15 us, 700 times per training step)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype in (torch.float32, torch.float16, torch.uint8, torch.int32), (t.device, t.dtype)
    assert t.is_contiguous(), "sos_b200 ops need contiguous tensors"
    return C.c_void_p(t.data_ptr())


def _count(n=1):
    _lib.launch_count += n


# ----------------------------------------------------------------------------------------------- per-kernel timing (bench.py)
_prof = None
_pool, _pool_next = [], 0


def profile_start(max_events=16384):
    """Start recording a CUDA event pair (on the launching stream) around every tensor-core / transform launch.  The events come
    from a pool created here, outside the timed region: constructing ~1000 event objects per step on the fly costs the host
    several ms per step, enough to make a GPU-bound step host-bound."""
    global _prof, _pool_next
    while len(_pool) < max_events:
        _pool.append(torch.cuda.Event(enable_timing=True))
    _pool_next = 0
    _prof = []


def _event():
    global _pool_next
    if _pool_next < len(_pool):
        e = _pool[_pool_next]
        _pool_next += 1
        return e
    return torch.cuda.Event(enable_timing=True)


def _pb():
    if _prof is None:
        return None
    e = _event()
    e.record()
    return e


def _pe(name, e0, flops=0.0, nbytes=0.0):
    if e0 is None:
        return
    e1 = _event()
    e1.record()
    _prof.append((name, e0, e1, flops, nbytes))


def profile_stop():
    """-> {kernel family: {"n": launches, "ms": device time, "flops": algorithmic FLOPs, "bytes": algorithmic bytes}}"""
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, fl, by in rec:
        d = out.setdefault(name, {"n": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["n"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += fl
        d["bytes"] += by
    return out


def init():
    if not torch.cuda.is_available():
        raise _lib.SosError("sos_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    check(lib().sos_init(), "sos_init")


def view8(H, W, Hp=None, Wp=None, ph=0, pw=0, ld=0, coff=0):
    Hp = H if Hp is None else Hp
    Wp = W if Wp is None else Wp
    return (_I32 * 8)(H, W, Hp, Wp, ph, pw, ld, coff)


# ----------------------------------------------------------------------------------------------- half-precision maps
# Activations between the tensor-core GEMMs are IEEE half NHWC arrays (same 11-bit significand as TF32, half the bytes, twice the
# MMA rate).  Autograd would cast a gradient to the dtype of the tensor it belongs to, and the gradients of these maps must stay
# fp32, so a half map travels through the graph behind an fp32-typed HANDLE of the same logical shape (N, H, W, C): strides
# (HWC/2, WC/2, C/2, 1) over a storage of N*H*W*C/2 + C/2 floats, i.e. exactly the half array's bytes.  Only the kernels below
# touch its memory, through `hv` (the real contiguous half tensor) or `fv` (the same bytes as fp32 pairs, for the copy kernels).
def new_half(shape, device):
    N, H, W, Cn = shape
    assert Cn % 8 == 0, "half maps need a multiple of 8 channels (16-byte pixels)"
    store = torch.empty(N * H * W * Cn // 2 + Cn // 2, device=device, dtype=torch.float32)
    return store.as_strided((N, H, W, Cn), (H * W * Cn // 2, W * Cn // 2, Cn // 2, 1))


def is_half_handle(t):
    return t.dtype == torch.float32 and t.dim() == 4 and t.shape[3] > 1 and t.stride(3) == 1 and t.stride(2) * 2 == t.shape[3]


def hv(h):
    assert is_half_handle(h), "expected a half-map handle (ops.new_half)"
    return torch.empty(0, device=h.device, dtype=torch.float16).set_(h.untyped_storage(), h.storage_offset() * 2, tuple(h.shape))


def fv(h):
    assert is_half_handle(h), "expected a half-map handle (ops.new_half)"
    N, H, W, Cn = h.shape
    return torch.empty(0, device=h.device, dtype=torch.float32).set_(h.untyped_storage(), h.storage_offset(), (N, H, W, Cn // 2))


def to_half(x, cd=None, scaled=False, out=None):
    """x (..., cs) fp32 -> real half tensor (..., cd) zero-padded; scaled: multiplied by the power of two that brings the RMS
    to ~1, returned as scal = [s, 1/s, sum x^2] (device floats; scal[1:2] is the out_scale of the consuming GEMMs)."""
    cs = x.shape[-1]
    cd = cs if cd is None else cd
    rows = x.numel() // cs
    if out is None:
        out = torch.empty(*x.shape[:-1], cd, device=x.device, dtype=torch.float16)
    scal = torch.zeros(3, device=x.device, dtype=torch.float32) if scaled else None
    check(lib().sos_to_half(_p(x), rows, cs, _p(out), cd, _p(scal), _stream()), "sos_to_half")
    _count(2 if scaled else 1)
    return (out, scal) if scaled else out


# ----------------------------------------------------------------------------------------------- transforms
def frame_lo_table(n_bits, ratio):
    """int(i * ratio) for i = 0..n_bits, with the reference's own float arithmetic (M2/tools.py:347-349)."""
    return [int(i * ratio) for i in range(n_bits + 1)]


_FLO = {}


def _frame_lo_device(n_bits, ratio, device):
    """The table on the device, cached: building it per call is a synchronous pageable host-to-device copy, which stalls the
    host until the stream drains (once per training step, so the host could never run ahead of the GPU)."""
    key = (int(n_bits), float(ratio), str(device))
    t = _FLO.get(key)
    if t is None:
        t = _FLO[key] = torch.tensor(frame_lo_table(n_bits, ratio), dtype=torch.int32, device=device)
    return t


def stft(wave, bits=None, ratio=None, gate_mode=0, fused_gate=False):
    """wave (B, L) -> (B, 2, 256, T).  bits (B, n_bits) uint8 (0 = silent) gates the waveform first.  By default the gate runs
    as its own elementwise launch (sos_gate_wave) in front of the tensor-core STFT; `fused_gate=True` evaluates the mask inside
    the STFT's frame builders instead (same result; the branchy mask logic makes that kernel ~7x slower)."""
    B, L = wave.shape
    T = 1 + L // 158
    out = torch.empty(B, 2, 256, T, device=wave.device, dtype=torch.float32)
    if gate_mode and not fused_gate:
        wave = gate_wave(wave, bits, ratio, gate_mode)
        gate_mode = 0
    e0 = _pb()
    if gate_mode:
        nb = bits.shape[1]
        flo = _frame_lo_device(nb, ratio, wave.device)
        check(lib().sos_stft_forward(_p(wave), B, L, _p(out), _p(bits), nb, _p(flo), float(ratio), gate_mode, _stream()), "sos_stft_forward")
    else:
        check(lib().sos_stft_forward(_p(wave), B, L, _p(out), None, 0, None, 0.0, 0, _stream()), "sos_stft_forward")
    _pe("stft", e0, 0.0, 4.0 * B * L + 4.0 * out.numel())
    _count()
    return out


_ISTFT_W3 = {}


def _istft_weight(device):
    """Weight operand of the inverse transform as a GEMM: rows n = 0..399 (window samples), K = 512 (re | im bins) =
    c_k w[n] cos / -sin(2 pi k (n + 55) / 510) / 510 (librosa.istft: irfft of the 510-point frame, Hann(400) centred in it; the
    imaginary parts of bins 0 and 255 are ignored), split [hi | hi | lo] once per device."""
    key = str(device)
    if key not in _ISTFT_W3:
        n = torch.arange(400, dtype=torch.float64)
        k = torch.arange(256, dtype=torch.float64)
        win = 0.5 - 0.5 * torch.cos(2 * torch.pi * n / 400)
        ph = 2 * torch.pi * torch.remainder((n[:, None] + 55) * k[None, :], 510) / 510
        coef = torch.full((256,), 2.0, dtype=torch.float64)
        coef[0] = coef[255] = 1.0
        wr = coef * win[:, None] * torch.cos(ph) / 510
        wi = -coef * win[:, None] * torch.sin(ph) / 510
        wi[:, 0] = wi[:, 255] = 0.0
        wt = torch.cat([wr, wi], dim=1).float().to(device).contiguous()           # (400, 512)
        _ISTFT_W3[key] = split_weight(wt)
    return _ISTFT_W3[key]


def istft(spec, crm=None, tensor_core=True):
    """spec (B, 2, 256, T) -> (B, 158 (T-1)); with crm the complex-ratio-mask recovery runs first.
    Default: the windowed inverse DFT of all frames is ONE fp32-grade tensor-core GEMM ((B T) x 512 spectrogram rows times the
    512 x 400 table, ops.gemm3) followed by the gather overlap-add kernel; tensor_core=False runs the CUDA-core frames kernel
    (sos_istft_forward, which can also fuse the cRM recovery into its spectrum load)."""
    B, _, F, T = spec.shape
    assert F == 256
    out = torch.empty(B, 158 * (T - 1), device=spec.device, dtype=torch.float32)
    e0 = _pb()
    if not tensor_core:
        ws = torch.empty(B * T * 400, device=spec.device, dtype=torch.float32)
        check(lib().sos_istft_forward(_p(spec), _p(crm), B, T, _p(ws), _p(out), _stream()), "sos_istft_forward")
        _count(2)
    else:
        if crm is not None:
            spec = icrm_forward(spec, crm)
        w3 = _istft_weight(spec.device)
        a2 = torch.empty(2, B * T, 512, device=spec.device, dtype=torch.float32)
        # clip b: (512, T) with T contiguous -> rows b T .. b T + T - 1, K = 512 (a transposing split, batched over the clips)
        check(lib().sos_split_tf32(_p(spec), T, 512, 512, 1, T, 0, _p(a2), 512, B * T * 512, 2, B, 512 * T, T * 512, _stream()), "sos_split_tf32")
        _count()
        frames = gemm3(a2, w3, 400, tag="istft_gemm")
        check(lib().sos_istft_ola(_p(frames), B, T, _p(out), _stream()), "sos_istft_ola")
        _count()
    _pe("istft", e0, 0.0, 4.0 * spec.numel() * (2 if crm is not None else 1) + 4.0 * out.numel())
    return out


# ----------------------------------------------------------------------------------------------- audio loading (librosa.load)
_RESAMPLE_FILTERS = {}


def _kaiser_best(ratio, device):
    """resampy's 'kaiser_best' interpolation table (64 zero crossings x 512 samples, rolloff 0.9475937167399596, Kaiser beta
    14.769656459379492; resampy/filters.py sinc_window), scaled by the ratio when down-sampling, and its forward differences."""
    import numpy as np
    from scipy.signal.windows import kaiser
    key = (float(ratio), str(device))
    if key not in _RESAMPLE_FILTERS:
        num_bits, num_zeros = 512, 64
        n = num_bits * num_zeros
        win = kaiser(2 * n + 1, 14.769656459379492)[n:] * 0.9475937167399596 * np.sinc(0.9475937167399596 * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
        if ratio < 1:
            win = win * ratio
        delta = np.zeros_like(win)
        delta[:-1] = np.diff(win)
        _RESAMPLE_FILTERS[key] = (torch.tensor(win, dtype=torch.float64, device=device), torch.tensor(delta, dtype=torch.float64, device=device), num_bits)
    return _RESAMPLE_FILTERS[key]


def resample(wave, sr_orig, sr_new):
    """wave (B, n) fp32 -> (B, int(n * sr_new / sr_orig)): resampy.resample(filter='kaiser_best'), the resampler behind
    librosa.load(path, sr=...) (M2/predict.py:303, M1/dataset.py:226)."""
    import numpy as np
    B, n = wave.shape
    ratio = float(sr_new) / float(sr_orig)
    n_out = int(n * ratio)
    if n_out < 1:
        raise _lib.SosError(f"resample: input of {n} samples is too short for the ratio {ratio}")
    win, delta, num_table = _kaiser_best(ratio, wave.device)
    # the reference advances its time register by repeated addition of 1 / ratio (float64): the same cumulative sum here
    treg = np.concatenate([[0.0], np.cumsum(np.full(n_out - 1, 1.0 / ratio))]) if n_out > 1 else np.zeros(1)
    treg = torch.tensor(treg, dtype=torch.float64, device=wave.device)
    out = torch.empty(B, n_out, device=wave.device, dtype=torch.float32)
    pd = lambda a: C.c_void_p(a.data_ptr())
    check(lib().sos_resample(_p(wave), B, n, _p(out), n_out, pd(win), pd(delta), win.numel(), num_table, ratio, pd(treg), _stream()), "sos_resample")
    _count()
    return out


def gate_wave(wave, bits, ratio, mode=1, want_mask=False):
    B, L = wave.shape
    nb = bits.shape[1]
    flo = _frame_lo_device(nb, ratio, wave.device)
    out = torch.empty_like(wave)
    mask = torch.empty_like(wave) if want_mask else None
    check(lib().sos_gate_wave(_p(wave), B, L, _p(bits), nb, _p(flo), float(ratio), mode, _p(out), _p(mask), _stream()), "sos_gate_wave")
    _count()
    return (out, mask) if want_mask else out


def icrm_forward(Y, crm, a=0.1, b=0.0):
    B = Y.shape[0]
    plane = Y[0, 0].numel()
    rec = torch.empty_like(Y)
    check(lib().sos_icrm_forward(_p(Y), _p(crm), _p(rec), B, plane, a, b, _stream()), "sos_icrm_forward")
    _count()
    return rec


def icrm_backward(Y, crm, grad_rec, a=0.1):
    B = Y.shape[0]
    plane = Y[0, 0].numel()
    g = torch.empty_like(crm)
    check(lib().sos_icrm_backward(_p(Y), _p(crm), _p(grad_rec), _p(g), B, plane, a, _stream()), "sos_icrm_backward")
    _count()
    return g


# ----------------------------------------------------------------------------------------------- training-item construction
def add_signals(signal, noise, snr_db, norm=0.5):
    """M2/tools.py:217-276 on the device: signal, noise (B, L), snr_db (B,) -> mixed, clean, full_noise (B, L)."""
    B, L = signal.shape
    assert noise.shape == signal.shape and snr_db.numel() == B
    mixed, clean, full = torch.empty_like(signal), torch.empty_like(signal), torch.empty_like(signal)
    check(lib().sos_add_signals(_p(signal), _p(noise), _p(snr_db), B, L, float(norm or 0.0), _p(mixed), _p(clean), _p(full), _stream()),
          "sos_add_signals")
    _count()
    return mixed, clean, full


def crm_forward(clean_spec, mixed_spec, a=0.1, b=0.0):
    """fast_cRM_sigmoid (M2/transform.py:130-138) on (B, 2, 256, T) spectrograms."""
    B = clean_spec.shape[0]
    plane = clean_spec[0, 0].numel()
    out = torch.empty_like(clean_spec)
    check(lib().sos_crm_forward(_p(clean_spec), _p(mixed_spec), _p(out), B, plane, a, b, _stream()), "sos_crm_forward")
    _count()
    return out


# ----------------------------------------------------------------------------------------------- losses / optimiser
def mse_fwd_bwd(pred, target, want_grad, grad_scale=1.0):
    n = pred.numel()
    loss = torch.zeros(1, device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if want_grad else None
    check(lib().sos_mse_fwd_bwd(_p(pred), _p(target), n, _p(loss), _p(grad), 2.0 * grad_scale / n, _stream()), "sos_mse_fwd_bwd")
    _count()
    return loss / n, grad


def bce_fwd_bwd(logits, labels, want_grad, grad_scale=1.0):
    n = logits.numel()
    loss = torch.zeros(1, device=logits.device, dtype=torch.float32)
    grad = torch.empty_like(logits) if want_grad else None
    check(lib().sos_bce_logits_fwd_bwd(_p(logits), _p(labels), n, _p(loss), _p(grad), grad_scale / n, _stream()), "sos_bce_logits_fwd_bwd")
    _count()
    return loss / n, grad


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    check(lib().sos_adam_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), lr, beta1, beta2, eps, step, grad_scale,
                              _stream()), "sos_adam_step")
    _count()


def adam_step_dev(param, grad, exp_avg, exp_avg_sq, state, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """Adam with the optimiser clock in device memory (state = [lr, step, ...]); advances the step by one."""
    check(lib().sos_adam_step_dev(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), _p(state), beta1, beta2, eps, grad_scale,
                                  _stream()), "sos_adam_step_dev")
    _count(2)


# ----------------------------------------------------------------------------------------------- BatchNorm + activation
ACT_NONE, ACT_RELU, ACT_PRELU, ACT_SIGMOID = 0, 1, 2, 3
ACT_ROUND_TF32 = 16          # OR into `act`: round the output to TF32 (it feeds a tensor-core GEMM)


def round_tf32_(x):
    """In-place round-to-nearest to TF32 (operands of the tcgen05 kind::tf32 GEMMs are otherwise truncated)."""
    check(lib().sos_round_tf32(_p(x), x.numel(), _stream()), "sos_round_tf32")
    _count()
    return x


def bn_train_forward(y, gamma, beta, running_mean, running_var, eps, momentum, act, slope, conv_partial=None):
    """y (..., C) dense NHWC rows.  Returns z and the tensors backward needs; updates running stats in place.
    conv_partial: (G, 2, C) per-warp sums / sums of squares already produced by the convolution's epilogue."""
    Cn = y.shape[-1]
    rows = y.numel() // Cn
    stats = torch.empty(4, Cn, device=y.device, dtype=torch.float32)          # mean, invstd, scale, shift
    sp = [C.c_void_p(stats[i].data_ptr()) for i in range(4)]
    if conv_partial is not None:
        assert conv_partial.shape[1:] == (2, Cn) and conv_partial.is_contiguous()
        check(lib().sos_bn_finalize_partial(_p(conv_partial), conv_partial.shape[0], rows, Cn, _p(gamma), _p(beta), eps, momentum,
                                            _p(running_mean), _p(running_var), sp[0], sp[1], sp[2], sp[3], _stream()), "sos_bn_finalize_partial")
        _count(-1)
    else:
        G = lib().sos_bn_partial_blocks(rows, Cn)
        if G <= 0:
            raise _lib.SosError(f"BatchNorm over {Cn} channels is not supported (need a multiple of 4, <= 1024)")
        partial = torch.empty(G * 3 * Cn, device=y.device, dtype=torch.float32)
        check(lib().sos_bn_stats(_p(y), rows, Cn, _p(partial), _stream()), "sos_bn_stats")
        check(lib().sos_bn_finalize(_p(partial), rows, Cn, _p(gamma), _p(beta), eps, momentum, _p(running_mean), _p(running_var),
                                    sp[0], sp[1], sp[2], sp[3], _stream()), "sos_bn_finalize")
    z = torch.empty_like(y)
    H, W = y.shape[-3], y.shape[-2]
    check(lib().sos_bn_act(_p(y), _p(z), view8(H, W, ld=Cn), rows, Cn, C.c_void_p(stats[2].data_ptr()), C.c_void_p(stats[3].data_ptr()),
                           act, _p(slope), _stream()), "sos_bn_act")
    _count(3)
    return z, stats


def bn_finalize_partial(conv_partial, rows, gamma, beta, running_mean, running_var, eps, momentum):
    """Batch statistics from the convolution epilogue's partial sums (G, 2, C) -> stats (4, C) = mean, invstd, scale, shift."""
    Cn = gamma.numel()
    assert conv_partial.shape[1:] == (2, Cn) and conv_partial.is_contiguous()
    stats = torch.empty(4, Cn, device=gamma.device, dtype=torch.float32)
    sp = [C.c_void_p(stats[i].data_ptr()) for i in range(4)]
    check(lib().sos_bn_finalize_partial(_p(conv_partial), conv_partial.shape[0], rows, Cn, _p(gamma), _p(beta), eps, momentum,
                                        _p(running_mean), _p(running_var), sp[0], sp[1], sp[2], sp[3], _stream()), "sos_bn_finalize_partial")
    _count()
    return stats


def _dt(t):
    return 1 if t.dtype == torch.float16 else 0


def bn_act_apply(y, stats, act, slope, half):
    """z = act(y * scale + shift) with the finalized stats: a half-map handle (half=True; y fp32 or half storage) or a dense fp32 map."""
    Cn = y.shape[-1]
    rows = y.numel() // Cn
    sc, sh = C.c_void_p(stats[2].data_ptr()), C.c_void_p(stats[3].data_ptr())
    if half:
        z = new_half(y.shape, y.device)
        e0 = _pb()
        check(lib().sos_bn_act_half(_p(y), _dt(y), _p(hv(z)), rows, Cn, sc, sh, act & 15, _p(slope), _stream()), "sos_bn_act_half")
        _pe("bn_fwd", e0, 0.0, y.numel() * (y.element_size() + 2.0))
    else:
        assert y.dtype == torch.float32
        z = torch.empty_like(y)
        check(lib().sos_bn_act(_p(y), _p(z), view8(y.shape[-3], y.shape[-2], ld=Cn), rows, Cn, sc, sh, act & 15, _p(slope), _stream()), "sos_bn_act")
    _count()
    return z


def bn_train_backward_half(dz, y, stats, act, slope, grad_into=None, dz_inv=None, pre_partial=None):
    """BatchNorm + activation backward with the gradient w.r.t. the conv output written as a scaled half operand.
    -> dy (real half tensor), dgamma, dbeta, dslope, scal [s, 1/s, sum dy^2].
    dz, y: fp32 or half storage; dz_inv: device scalar, the inverse of the power-of-two scale a half dz still carries.
    grad_into = (gamma.grad, beta.grad, slope.grad or None): the parameter gradients are ADDED straight into these (their first
    gamma.grad.numel() channels; the map may carry zero-padded channels) and None is returned in their place.
    pre_partial (rows, 4, C): pass 1 (the reduction over dz, y) already came out of the epilogue of the GEMM that produced dz
    (conv_tc(bnr=...)): only the finalize and apply kernels run."""
    Cn = y.shape[-1]
    rows = y.numel() // Cn
    G = lib().sos_bn_partial_blocks(rows, Cn)
    partial = torch.empty(G * 4 * Cn, device=y.device, dtype=torch.float32)
    dy = torch.empty(y.shape, device=y.device, dtype=torch.float16)
    out = torch.empty(4, Cn, device=y.device, dtype=torch.float32)             # dgamma, dbeta, m1, m2
    prelu = (act & 15) == ACT_PRELU
    # (sum dy^2 accumulates in scal[2]: zeroed by the reduction kernel itself, by a fill only when that pass is skipped)
    scal = (torch.zeros if pre_partial is not None else torch.empty)(3, device=y.device, dtype=torch.float32)
    sp = lambda i: C.c_void_p(stats[i].data_ptr())
    op = lambda i: C.c_void_p(out[i].data_ptr())
    nbytes = y.numel() * (2.0 * dz.element_size() + 2.0 * y.element_size() + 2.0)
    e0 = _pb()
    if pre_partial is not None:
        assert (act & 15) == ACT_RELU and dz.dtype == torch.float16 and y.dtype == torch.float16 and pre_partial.shape[1:] == (4, Cn)
        acc = grad_into is not None
        gg, gb = (grad_into[0], grad_into[1]) if acc else (out[0], out[1])
        nbytes = y.numel() * 6.0
        check(lib().sos_bn_act_backward_half_pre(_p(dz), _dt(dz), _p(dz_inv), _p(y), _dt(y), _p(dy), rows, Cn, sp(2), sp(3), sp(0), sp(1), act & 15,
                                                 None, _p(pre_partial), pre_partial.shape[0], _p(gg), _p(gb), None, op(2), op(3), _p(scal),
                                                 1 if acc else 0, gg.numel() if acc else 0, _stream()), "sos_bn_act_backward_half_pre")
        _pe("bn_bwd", e0, 0.0, nbytes)
        _count(2)
        return (dy, None, None, None, scal) if acc else (dy, out[0], out[1], None, scal)
    if grad_into is not None:
        gg, gb, gs = grad_into
        assert gg.is_contiguous() and gb.is_contiguous() and gg.numel() == gb.numel() <= Cn and (not prelu or gs is not None)
        check(lib().sos_bn_act_backward_half(_p(dz), _dt(dz), _p(dz_inv), _p(y), _dt(y), _p(dy), rows, Cn, sp(2), sp(3), sp(0), sp(1), act & 15,
                                             _p(slope), _p(partial), _p(gg), _p(gb), _p(gs) if prelu else None, op(2), op(3), _p(scal), 1,
                                             gg.numel(), _stream()), "sos_bn_act_backward_half")
        _pe("bn_bwd", e0, 0.0, nbytes)
        _count(3)
        return dy, None, None, None, scal
    dslope = torch.zeros(1, device=y.device, dtype=torch.float32) if prelu else None
    check(lib().sos_bn_act_backward_half(_p(dz), _dt(dz), _p(dz_inv), _p(y), _dt(y), _p(dy), rows, Cn, sp(2), sp(3), sp(0), sp(1), act & 15,
                                         _p(slope), _p(partial), op(0), op(1), _p(dslope), op(2), op(3), _p(scal), 0, 0, _stream()),
          "sos_bn_act_backward_half")
    _pe("bn_bwd", e0, 0.0, nbytes)
    _count(3)
    return dy, out[0], out[1], dslope, scal


def bn_train_backward(dz, y, stats, act, slope):
    Cn = y.shape[-1]
    rows = y.numel() // Cn
    G = lib().sos_bn_partial_blocks(rows, Cn)
    partial = torch.empty(G * 3 * Cn, device=y.device, dtype=torch.float32)
    dy = torch.empty_like(y)
    out = torch.empty(4, Cn, device=y.device, dtype=torch.float32)             # dgamma, dbeta, m1, m2
    dslope = torch.zeros(1, device=y.device, dtype=torch.float32) if (act & 15) == ACT_PRELU else None
    H, W = y.shape[-3], y.shape[-2]
    sp = lambda i: C.c_void_p(stats[i].data_ptr())
    op = lambda i: C.c_void_p(out[i].data_ptr())
    check(lib().sos_bn_act_backward(_p(dz), view8(H, W, ld=Cn), _p(y), _p(dy), rows, Cn, sp(2), sp(3), sp(0), sp(1), act, _p(slope),
                                    _p(partial), op(0), op(1), _p(dslope), op(2), op(3), _stream()), "sos_bn_act_backward")
    _count(3)
    return dy, out[0], out[1], dslope


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps):
    Cn = gamma.numel()
    coeffs = torch.empty(2, Cn, device=gamma.device, dtype=torch.float32)
    check(lib().sos_bn_eval_coeffs(Cn, _p(gamma), _p(beta), _p(running_mean), _p(running_var), eps, C.c_void_p(coeffs[0].data_ptr()),
                                   C.c_void_p(coeffs[1].data_ptr()), _stream()), "sos_bn_eval_coeffs")
    _count()
    return coeffs[0], coeffs[1]


# ----------------------------------------------------------------------------------------------- layout
def nchw_to_nhwc(x, channels_padded):
    """(B, C, H, W) -> dense (B, H, W, Cp) with zero padded channels."""
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, channels_padded, device=x.device, dtype=torch.float32)
    check(lib().sos_nchw_to_nhwc(_p(x), B, Cc, _p(out), view8(H, W, ld=channels_padded), channels_padded, _stream()), "sos_nchw_to_nhwc")
    _count()
    return out


def nchw_to_nhwc_half(x, channels_padded):
    """(B, C, H, W) fp32 -> half-map handle (B, H, W, Cp) with zero padded channels, one pass."""
    B, Cc, H, W = x.shape
    z = new_half((B, H, W, channels_padded), x.device)
    check(lib().sos_nchw_to_nhwc_half(_p(x), B, Cc, H, W, _p(hv(z)), channels_padded, _stream()), "sos_nchw_to_nhwc_half")
    _count()
    return z


def accumulate_wgrad(dwt, grad, clear=False):
    """grad (R, Cc, kh, kw) += dwt (ntaps, RP, CP) (the weight-gradient kernel's tap-major buffer, padded rows / columns);
    clear: dwt is zeroed again on the way (a `wgrad_workspace`)."""
    ntaps, RP, CP = dwt.shape
    R, Cc = grad.shape[0], grad.shape[1]
    assert grad.is_contiguous() and grad.shape[2] * grad.shape[3] == ntaps and R <= RP and Cc <= CP
    fn = lib().sos_accumulate_wgrad_clear if clear else lib().sos_accumulate_wgrad
    check(fn(_p(dwt), ntaps, RP, CP, R, Cc, _p(grad), _stream()), "sos_accumulate_wgrad")
    _count()


_WGRAD_WS = {}


def wgrad_workspace(ntaps, rows, cols, device):
    """A persistent zeroed (ntaps, rows, cols) fp32 buffer for conv_wgrad(dw=...) + accumulate_wgrad(clear=True): every user leaves it
    zeroed, all users run in order on one stream (a weight-gradient side stream; the buffer is per stream), so one buffer per shape
    serves every layer and a training step needs no fill launches for its weight gradients."""
    key = (str(device), torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0, ntaps, rows, cols)
    ws = _WGRAD_WS.get(key)
    if ws is None:
        ws = _WGRAD_WS[key] = torch.zeros(ntaps, rows, cols, device=device, dtype=torch.float32)
    return ws


def nhwc_to_nchw(x, channels):
    B, H, W, Cp = x.shape
    out = torch.empty(B, channels, H, W, device=x.device, dtype=torch.float32)
    check(lib().sos_nhwc_to_nchw(_p(x), view8(H, W, ld=Cp), B, channels, _p(out), _stream()), "sos_nhwc_to_nchw")
    _count()
    return out


def copy_view(src, sview, dst, dview, batch, channels, accumulate=False):
    check(lib().sos_copy_view(_p(src), sview, _p(dst), dview, batch, channels, int(accumulate), _stream()), "sos_copy_view")
    _count()


def copy_view_backward(gdst, dview, gsrc, sview, batch, channels):
    check(lib().sos_copy_view_backward(_p(gdst), dview, _p(gsrc), sview, batch, channels, _stream()), "sos_copy_view_backward")
    _count()


def copy_view_fold(gpad, pview, dst, dview, batch, channels):
    """dst window = interior of the reflect-padded gradient map + the border positions mirrored onto it (gpad is not modified)."""
    check(lib().sos_copy_view_fold(_p(gpad), pview, _p(dst), dview, batch, channels, _stream()), "sos_copy_view_fold")
    _count()


def reflect_fill(buf, H, W, pad):
    B, Hp, Wp, Cn = buf.shape
    assert Hp == H + 2 * pad and Wp == W + 2 * pad
    check(lib().sos_reflect_fill(_p(buf), B, H, W, pad, Cn, _stream()), "sos_reflect_fill")
    _count()


def reflect_fold(gbuf, H, W, pad):
    B, Hp, Wp, Cn = gbuf.shape
    check(lib().sos_reflect_fold(_p(gbuf), B, H, W, pad, Cn, _stream()), "sos_reflect_fold")
    _count()


def feat_to_seq(x, out, V, coff):
    """x NHWC (B, F, T, C) -> out (V, B, ld)[:, :, coff + c*F + f] with nearest resample T -> V."""
    B, F, T, Cn = x.shape
    check(lib().sos_feat_to_seq(_p(x), B, F, T, Cn, _p(out), V, out.shape[2], coff, _stream()), "sos_feat_to_seq")
    _count()


def feat_to_seq_backward(gout, shape, V, coff):
    B, F, T, Cn = shape
    gin = torch.zeros(shape, device=gout.device, dtype=torch.float32)
    check(lib().sos_feat_to_seq_backward(_p(gout), B, F, T, Cn, _p(gin), V, gout.shape[2], coff, _stream()), "sos_feat_to_seq_backward")
    _count()
    return gin


# ----------------------------------------------------------------------------------------------- tensor-core tap GEMM
def _i32arr(v):
    return (_I32 * len(v))(*[int(a) for a in v])


def pack_taps(w, taps, k_padded, round_tf32=True):
    """w: 4-D weight view (rows, K, kh, kw) with arbitrary strides over one storage -> (rows, len(taps)*k_padded) GEMM operand,
    column t*k_padded + k = w[r, k, taps[t][0], taps[t][1]] (zero padded, TF32-rounded): one launch."""
    assert w.dim() == 4 and w.is_cuda and w.dtype == torch.float32
    R, K = w.shape[0], w.shape[1]
    st = w.stride()
    out = torch.empty(R, len(taps) * k_padded, device=w.device, dtype=torch.float32)
    offs = _i32arr([a * st[2] + b * st[3] for a, b in taps])
    check(lib().sos_pack_taps(C.c_void_p(w.data_ptr()), R, K, k_padded, st[0], st[1], len(taps), offs, int(round_tf32), _p(out), _stream()),
          "sos_pack_taps")
    _count()
    return out


def pack_desc(w, taps, k_padded, out):
    """The sos_pack_desc of pack_taps_half(w, taps, k_padded) -> out (for pack_taps_half_multi)."""
    from ._lib import PackDesc
    st = w.stride()
    d = PackDesc()
    d.w, d.out_half = w.data_ptr(), out.data_ptr()
    d.rows, d.K, d.KP, d.row_stride, d.k_stride, d.ntaps = w.shape[0], w.shape[1], k_padded, st[0], st[1], len(taps)
    for i, t in enumerate(taps):
        d.tap_off[i] = -1 if t is None else t[0] * st[2] + t[1] * st[3]         # None: a zero tap (padding of a folded operand)
    return d


def im2col_half(x, tap_dh, tap_dw, OH, OW, Kc):
    """x (N, H, W, Cp) half, its first TWO channels real -> (N, OH, OW, Kc) half with the taps folded into the channel axis:
    column 2 t + c = x[n, oh + dh_t, ow + dw_t, c] (zero outside the image, zero padding columns)."""
    N, H, W, Cp = x.shape
    assert x.dtype == torch.float16 and x.is_contiguous() and Kc % 8 == 0 and 2 * len(tap_dh) <= Kc
    out = torch.empty(N, OH, OW, Kc, device=x.device, dtype=torch.float16)
    check(lib().sos_im2col_half(_p(x), N, H, W, Cp, len(tap_dh), _i32arr(tap_dh), _i32arr(tap_dw), OH, OW, _p(out), Kc, _stream()), "sos_im2col_half")
    _count()
    return out


def pack_table(descs, device):
    """Device copy of a list of sos_pack_desc (a uint8 tensor)."""
    raw = b"".join(bytes(d) for d in descs)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


def pack_taps_half_multi(table, n):
    check(lib().sos_pack_taps_half_multi(C.c_void_p(table.data_ptr()), n, _stream()), "sos_pack_taps_half_multi")
    _count()


def pack_taps_half(w, taps, k_padded, out=None):
    """pack_taps with a half output (operand of the kind::f16 GEMMs)."""
    assert w.dim() == 4 and w.is_cuda and w.dtype == torch.float32
    R, K = w.shape[0], w.shape[1]
    st = w.stride()
    if out is None:
        out = torch.empty(R, len(taps) * k_padded, device=w.device, dtype=torch.float16)
    offs = _i32arr([-1 if t is None else t[0] * st[2] + t[1] * st[3] for t in taps])
    check(lib().sos_pack_taps_half(C.c_void_p(w.data_ptr()), R, K, k_padded, st[0], st[1], len(taps), offs, _p(out), _stream()),
          "sos_pack_taps_half")
    _count()
    return out


def conv_tc(x, wk, tap_dh, tap_dw, Cout, OH, OW, stride=1, y=None, y_coff=0, lattice=(1, 1, 0, 0), epi_scale=None, epi_shift=None,
            act=0, slope=None, force_plan=-1, plan_out=None, k_real=None, tag="conv_fwd", want_stats=False, y_half=False, out_scale=None,
            bnr=None):
    """Tap-list implicit GEMM on tcgen05 (see include/sos_b200.h: sos_conv2d_tc).

    x (N, H, W, Cin) NHWC; wk (Cout, ntaps*Cin); y (N, YH, YW, Cy) is allocated when None (dense, Cy = Cout
    rounded up to 8, zero filled if padded).  x and wk are both fp32 (read as TF32) or both half (kind::f16); y is fp32, or a
    real half tensor with y_half.  out_scale: device scalar multiplied into the outputs.
    bnr = (y_below, stats_below): a data-gradient call whose output is the dz of the layer below asks the epilogue for that layer's
    BatchNorm-backward reduction as well; returns (y, partial or None) -- None when the kernel serving the call cannot do it."""
    N, H, W, Cin = x.shape
    ntaps = len(tap_dh)
    assert wk.shape == (Cout, ntaps * Cin), (wk.shape, Cout, ntaps, Cin)
    if y is None:
        Cy = (Cout + 7) // 8 * 8
        osh, osw, _, _ = lattice
        alloc = torch.zeros if (Cy != Cout) else torch.empty
        y = alloc(N, OH * osh, OW * osw, Cy, device=x.device, dtype=torch.float16 if y_half else torch.float32)
    assert x.dtype == wk.dtype and x.dtype in (torch.float32, torch.float16)
    a = ConvArgs()
    a.x_dtype = 1 if x.dtype == torch.float16 else 0
    a.y_dtype = 1 if y.dtype == torch.float16 else 0
    a.out_scale = out_scale.data_ptr() if out_scale is not None else None
    a.x, a.wk, a.y = x.data_ptr(), wk.data_ptr(), y.data_ptr()
    dh, dw = _i32arr(tap_dh), _i32arr(tap_dw)
    a.tap_dh, a.tap_dw = dh, dw
    a.N, a.H, a.W, a.Cin = N, H, W, Cin
    a.Cout, a.OH, a.OW = Cout, OH, OW
    a.ntaps, a.stride = ntaps, stride
    a.YH, a.YW, a.Cy, a.y_coff = y.shape[1], y.shape[2], y.shape[3], y_coff
    a.osh, a.osw, a.oph, a.opw = lattice
    a.epi_scale = epi_scale.data_ptr() if epi_scale is not None else None
    a.epi_shift = epi_shift.data_ptr() if epi_shift is not None else None
    a.act = act
    a.slope = slope.data_ptr() if slope is not None else None
    a.force_plan = force_plan
    po = (_I32 * 8)() if (plan_out is not None or _prof is not None) else None
    a.plan_out = po
    partial, rows = None, None
    if want_stats:                                   # BatchNorm partial sums of the raw outputs, written by the epilogue
        partial = torch.empty(lib().sos_conv_stats_rows(), 2, y.shape[3], device=x.device, dtype=torch.float32)
        rows = (_I32 * 1)()
        a.stats_partial, a.stats_channels, a.stats_rows_out = partial.data_ptr(), y.shape[3], rows
    bpart = brows = None
    if bnr is not None:
        yb, sb = bnr
        assert yb.dtype == torch.float16 and yb.shape == y.shape and yb.is_contiguous() and y_coff == 0
        bpart = torch.empty(lib().sos_conv_stats_rows(), 4, y.shape[3], device=x.device, dtype=torch.float32)
        brows = (_I32 * 1)()
        a.bnr_y, a.bnr_partial, a.bnr_channels, a.bnr_rows_out = yb.data_ptr(), bpart.data_ptr(), y.shape[3], brows
        a.bnr_scale, a.bnr_shift, a.bnr_mean, a.bnr_invstd = sb[2].data_ptr(), sb[3].data_ptr(), sb[0].data_ptr(), sb[1].data_ptr()
    assert x.is_contiguous() and wk.is_contiguous() and y.is_contiguous()
    e0 = _pb()
    check(lib().sos_conv2d_tc(C.byref(a), _stream()), "sos_conv2d_tc")
    # (profiling: launches served by the row-streaming kernel -- plan_out[0] == 2 -- are booked as their own family)
    _pe(tag + ("_row" if (po is not None and po[0] == 2 and e0 is not None) else ""), e0, 2.0 * N * OH * OW * Cout * (k_real or Cin) * ntaps)
    _count()
    if plan_out is not None:
        plan_out[:] = list(po)
    if bnr is not None:
        return y, (bpart[:brows[0]] if brows[0] > 0 else None)
    if want_stats:
        return y, partial[:rows[0]]
    return y


def conv_wgrad(x, dy, tap_dh, tap_dw, Cout, OH, OW, stride=1, dy_coff=0, force_plan=-1, plan_out=None, real=None, out_scale=None, workspace=False):
    """dw[t][co][ci] = sum_pixels dy[p][dy_coff+co] * x[p*stride + off_t][ci]  ->  (ntaps, Cout, Cin); workspace: the sums go into the
    shared `wgrad_workspace` of that shape (which the caller must empty with accumulate_wgrad(clear=True))."""
    N, H, W, Cin = x.shape
    ntaps = len(tap_dh)
    assert dy.shape[0] == N and dy.shape[1] == OH and dy.shape[2] == OW, (dy.shape, N, OH, OW)
    dw = wgrad_workspace(ntaps, Cout, Cin, x.device) if workspace else torch.zeros(ntaps, Cout, Cin, device=x.device, dtype=torch.float32)
    assert x.dtype == dy.dtype and x.dtype in (torch.float32, torch.float16)
    a = WgradArgs()
    a.dtype = 1 if x.dtype == torch.float16 else 0
    a.out_scale = out_scale.data_ptr() if out_scale is not None else None
    a.x, a.dy, a.dw = x.data_ptr(), dy.data_ptr(), dw.data_ptr()
    dh, dwv = _i32arr(tap_dh), _i32arr(tap_dw)
    a.tap_dh, a.tap_dw = dh, dwv
    a.N, a.H, a.W, a.Cin = N, H, W, Cin
    a.Cout, a.OH, a.OW, a.Cdy, a.dy_coff = Cout, OH, OW, dy.shape[3], dy_coff
    a.ntaps, a.stride = ntaps, stride
    a.force_plan = force_plan
    po = (_I32 * 8)() if plan_out is not None else None
    a.plan_out = po
    assert x.is_contiguous() and dy.is_contiguous()
    e0 = _pb()
    check(lib().sos_conv2d_wgrad(C.byref(a), _stream()), "sos_conv2d_wgrad")
    ci_r, co_r = real or (Cin, Cout)
    _pe("conv_wgrad", e0, 2.0 * N * OH * OW * co_r * ci_r * ntaps)
    _count()
    if plan_out is not None:
        plan_out[:] = list(po)
    return dw


# ----------------------------------------------------------------------------------------------- fp32-grade GEMMs on the tap GEMM
# The LSTM input projections and the MLP heads (M1/networks.py:95-98, M2/networks.py:64-70) are contractions over 200..6500 fp32
# values whose results feed a sigmoid mask: they run at fp32-grade accuracy on the tensor cores as ONE 3-tap launch of the TF32 tap
# GEMM over split operands (x = hi + lo, both TF32-exact; a @ b^T = hi hi + lo hi + hi lo, error ~2^-21): the activation operand is
# the stack [hi; lo] (N = 1, H = 2, W = rows, C = K), tap t reads stack row {0, 1, 0}[t], and the weight operand carries the slots
# [hi | hi | lo] along K.
def _r8(n):
    return (n + 7) // 8 * 8


def _split_into(src, transpose, k_shift, out, ld_out, slot_stride, n_slots, col_offset=0):
    """Rows of the operand = rows of src (or its columns with `transpose`), K axis = the other one; writes slots of width KP = K
    rounded up to 8 at column `col_offset` of rows of length ld_out."""
    assert src.dim() == 2 and src.stride(1) == 1 and src.dtype == torch.float32 and src.is_cuda
    R, K = src.shape
    dst = C.c_void_p(out.data_ptr() + 4 * col_offset)
    if not transpose:
        check(lib().sos_split_tf32(_p_any(src), R, K, _r8(K), src.stride(0), 1, 0, dst, ld_out, slot_stride, n_slots, 1, 0, 0, _stream()), "sos_split_tf32")
    else:
        check(lib().sos_split_tf32(_p_any(src), K, R, _r8(R), 1, src.stride(0), k_shift, dst, ld_out, slot_stride, n_slots, 1, 0, 0, _stream()),
              "sos_split_tf32")
    _count()


def split_act(x, transpose=False, k_shift=0, passes=3):
    """x (R, K) fp32 (any row stride, unit column stride) -> activation stack (2, R, KP) [hi; lo], KP = K rounded up to 8.
    transpose: the operand is x^T, i.e. the stack is (2, K, RP) (RP = R rounded up to 8); k_shift then shifts along R:
    element [k][r] = x[r + k_shift][k] (zero outside), which expresses h_{t-1} / h_{t+1} of an LSTM output.
    passes = 1: only the TF32-rounded part, (1, rows, KP): the operand of a single-pass TF32 GEMM (gradient GEMMs of the half mode)."""
    R, K = x.shape
    rows, kp = (K, _r8(R)) if transpose else (R, _r8(K))
    ns = 2 if passes == 3 else 1
    out = torch.empty(ns, rows, kp, device=x.device, dtype=torch.float32)
    _split_into(x, transpose, k_shift, out, kp, rows * kp, ns)
    return out


def split_weight(w, transpose=False, k_shift=0, passes=3):
    """w (Cout, K) fp32 (unit column stride) -> weight operand (Cout, 3 * KP) [hi | hi | lo]; transpose: the operand is w^T
    (K, 3 * RP) with the same k_shift semantics as split_act.  passes = 1: (rows, KP), the TF32-rounded part only."""
    R, K = w.shape
    rows, kp = (K, _r8(R)) if transpose else (R, _r8(K))
    out = torch.empty(rows, passes * kp, device=w.device, dtype=torch.float32)
    _split_into(w, transpose, k_shift, out, passes * kp, kp, passes)
    return out


def split_weight_cat_t(ws, passes=3):
    """Weight operand of x @ [w_0; w_1; ...] (the w_i (R_i, K) stacked along their rows, which is the contraction axis here):
    (K, 3 * sum RP_i) with slot s holding [w_0^T | w_1^T | ...] (passes = 1: one slot)."""
    K = ws[0].shape[1]
    kp = sum(_r8(w.shape[0]) for w in ws)
    out = torch.empty(K, passes * kp, device=ws[0].device, dtype=torch.float32)
    off = 0
    for w in ws:
        _split_into(w, True, 0, out, passes * kp, kp, passes, col_offset=off)
        off += _r8(w.shape[0])
    return out


def _p_any(t):
    assert t.is_cuda and t.dtype == torch.float32
    return C.c_void_p(t.data_ptr())


def gemm3(a2, w3, n_out, bias=None, act=0, tag="gemm", out=None, col=0):
    """a2 (2, M, KP) activation stack, w3 (n_out, 3 * KP) weight operand -> (M, n_out) fp32 = act(a @ w^T + bias); act 0 / 1 relu /
    3 sigmoid.  out (M, ld) + col: write into columns [col, col + n_out) of an existing buffer (col, ld multiples of 4).  A fresh
    output is padded to a multiple of 4 columns; the returned view drops the padding."""
    ns, M, KP = a2.shape
    one = ns == 1                                     # single-pass TF32 (hi . hi): one tap over the TF32-rounded operands
    assert w3.shape == (n_out, (1 if one else 3) * KP), (w3.shape, n_out, KP)
    if out is None:
        Cy = (n_out + 3) // 4 * 4
        out = torch.empty(M, Cy, device=a2.device, dtype=torch.float32)
    Cy = out.shape[1]
    conv_tc(a2.view(1, ns, M, KP), w3, [0] if one else [0, 1, 0], [0] if one else [0, 0, 0], n_out, 1, M, 1, y=out.view(1, 1, M, Cy), y_coff=col,
            epi_shift=bias, act=act, force_plan=1, tag=tag, k_real=KP)
    return out if (Cy == n_out and col == 0) else out[:, col:col + n_out]


def axpy_(dst, src, alpha=1.0, base=None):
    """dst = (base if given else dst) + alpha * src."""
    assert dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel() and (base is None or base.is_contiguous())
    check(lib().sos_axpy(_p(dst), _p(src), dst.numel(), alpha, _p(base), _stream()), "sos_axpy")
    _count()
    return dst


def seq_to_map(h, channels):
    """h (T, B, ld) sequence rows -> (B, channels, T) maps (the mask head's permute + view, M2/networks.py:92-93)."""
    T, B, ld = h.shape
    out = torch.empty(B, channels, T, device=h.device, dtype=torch.float32)
    check(lib().sos_seq_map(_p(h), _p(out), T, B, channels, ld, 1, _stream()), "sos_seq_map")
    _count()
    return out


def map_to_seq(g, T, B, channels):
    """adjoint of seq_to_map: g (B, channels, T) -> (T, B, channels)."""
    out = torch.empty(T, B, channels, device=g.device, dtype=torch.float32)
    check(lib().sos_seq_map(_p(g), _p(out), T, B, channels, channels, 0, _stream()), "sos_seq_map")
    _count()
    return out


def bias_act_backward(dy, y, act, dbias=None, want_dpre=True):
    """dpre = dy * act'(y) for y = act(pre) (act 0 / 1 relu / 3 sigmoid), dbias += column sums of dpre.  dy, y (rows, cols), unit
    column stride, same row stride; want_dpre=False: only the column sums."""
    rows, cols = dy.shape
    assert dy.stride(1) == 1 and y.stride(1) == 1 and dy.stride(0) == y.stride(0)
    dpre = torch.empty_strided(dy.shape, dy.stride(), device=dy.device, dtype=torch.float32) if want_dpre else None
    check(lib().sos_bias_act_backward(_p_any(dy), _p_any(y), _p_any(dpre) if want_dpre else None, rows, cols, dy.stride(0), act, _p(dbias),
                                      _stream()), "sos_bias_act_backward")
    _count()
    return dpre


# ----------------------------------------------------------------------------------------------- LSTM recurrence
def lstm_forward(gx, w_hh):
    """gx (T, B, 2, 4H), w_hh (2, 4H, H) -> out (T, B, 2H), gates (T, B, 2, 4H), cell (T, B, 2, H)."""
    T, B, _, H4 = gx.shape
    H = H4 // 4
    out = torch.empty(T, B, 2 * H, device=gx.device, dtype=torch.float32)
    gates = torch.empty_like(gx)
    cell = torch.empty(T, B, 2, H, device=gx.device, dtype=torch.float32)
    e0 = _pb()
    check(lib().sos_lstm_forward(_p(gx), _p(w_hh), T, B, H, _p(out), _p(gates), _p(cell), _stream()), "sos_lstm_forward")
    _pe("lstm_fwd", e0)
    _count()
    return out, gates, cell


def lstm_backward(dout, w_hh, out, gates, cell):
    T, B, _, H4 = gates.shape
    H = H4 // 4
    dgx = torch.empty_like(gates)
    dc = torch.empty(B, 2, H, device=dout.device, dtype=torch.float32)
    e0 = _pb()
    check(lib().sos_lstm_backward(_p(dout), _p(w_hh), _p(out), _p(gates), _p(cell), T, B, H, _p(dgx), None, _p(dc), _stream()),
          "sos_lstm_backward")
    _pe("lstm_bwd", e0)
    _count()
    return dgx
