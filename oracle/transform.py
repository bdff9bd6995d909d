"""numpy restatement of the reference's signal transforms (TEST INFRASTRUCTURE).

Follows M2/transform.py (= M1/transform.py, byte identical) of the reference:
  fast_stft            transform.py:188-193  -> librosa.stft(data, 510, 158, 400)
  fast_istft           transform.py:196-202  -> librosa.istft(S, 158, 400)
  cRM_sigmoid_*        transform.py:92-99
  generate_cRM         transform.py:36-54
  fast_icRM_sigmoid    transform.py:141-153
  batch_fast_icRM_sigmoid  transform.py:156-169
librosa==0.7.1 (requirements.txt:4) is a third-party dependency that is not in
/root/reference; its published stft/istft algorithm (librosa/core/spectrum.py
of the 0.7.1 release) is restated below in float64 numpy.
"""
import numpy as np

N_FFT = 510          # transform.py:6
HOP_LENGTH = 158     # transform.py:7
WIN_LENGTH = 400     # transform.py:8


def hann_periodic(win_length=WIN_LENGTH):
    # scipy.signal.get_window('hann', n, fftbins=True)
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def padded_window(n_fft=N_FFT, win_length=WIN_LENGTH):
    # librosa.util.pad_center(window, n_fft)
    w = np.zeros(n_fft, dtype=np.float64)
    lpad = (n_fft - win_length) // 2
    w[lpad:lpad + win_length] = hann_periodic(win_length)
    return w


def num_frames(length, hop_length=HOP_LENGTH):
    return 1 + length // hop_length


def librosa_stft(y, n_fft=N_FFT, hop_length=HOP_LENGTH, win_length=WIN_LENGTH):
    """librosa 0.7.1 stft(center=True, pad_mode='reflect', dtype=complex64)."""
    y = np.asarray(y)
    w = padded_window(n_fft, win_length)
    yp = np.pad(y.astype(np.float64), n_fft // 2, mode="reflect")
    t = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(t)[None, :]
    frames = yp[idx] * w[:, None]
    return np.fft.rfft(frames, axis=0).astype(np.complex64)       # (1+n_fft/2, T)


def window_sumsquare(n_frames, n_fft=N_FFT, hop_length=HOP_LENGTH, win_length=WIN_LENGTH):
    # librosa.filters.window_sumsquare(norm=None, dtype=float32)
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=np.float32)
    win_sq = padded_window(n_fft, win_length) ** 2
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))].astype(np.float32)
    return x


def librosa_istft(S, hop_length=HOP_LENGTH, win_length=WIN_LENGTH):
    """librosa 0.7.1 istft(center=True, length=None, dtype=float32)."""
    S = np.asarray(S)
    n_fft = 2 * (S.shape[0] - 1)
    t = S.shape[1]
    w = padded_window(n_fft, win_length)
    y = np.zeros(n_fft + hop_length * (t - 1), dtype=np.float32)
    ytmp = w[:, None] * np.fft.irfft(S.astype(np.complex128), n=n_fft, axis=0)
    for i in range(t):
        y[i * hop_length:i * hop_length + n_fft] += ytmp[:, i].astype(np.float32)
    wss = window_sumsquare(t, n_fft, hop_length, win_length)
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:-(n_fft // 2)]


def real_imag_expand(c):            # transform.py:10-22 (dim='new')
    d = np.zeros((c.shape[0], c.shape[1], 2))
    d[:, :, 0] = np.real(c)
    d[:, :, 1] = np.imag(c)
    return d


def real_imag_shrink(f):            # transform.py:25-33 (dim='new')
    return f[:, :, 0] + f[:, :, 1] * 1j


def fast_stft(data, n_fft=N_FFT, hop_length=HOP_LENGTH, win_length=WIN_LENGTH):
    return real_imag_expand(librosa_stft(data, n_fft, hop_length, win_length))


def fast_istft(F, hop_length=HOP_LENGTH, win_length=WIN_LENGTH):
    return librosa_istft(real_imag_shrink(F), hop_length, win_length)


def generate_cRM(Y, S, eps=1e-8):   # transform.py:36-54 ; Y mixed, S clean, (F,T,2)
    den = Y[..., 0] ** 2 + Y[..., 1] ** 2 + eps
    M = np.zeros(Y.shape)
    M[..., 0] = (Y[..., 0] * S[..., 0] + Y[..., 1] * S[..., 1]) / den
    M[..., 1] = (Y[..., 0] * S[..., 1] - Y[..., 1] * S[..., 0]) / den
    return M


def cRM_sigmoid_compress(M, a=0.1, b=0):      # transform.py:92-94
    return 1.0 / (1.0 + np.exp(-a * M + b))


def cRM_sigmoid_recover(O, a=0.1, b=0):       # transform.py:97-99
    return 1.0 / a * (np.log(O / (1 - O + 1e-8) + 1e-10) + b)


def fast_cRM_sigmoid(Fclean, Fmix):           # transform.py:130-138
    return cRM_sigmoid_compress(generate_cRM(Fmix, Fclean))


def fast_icRM_sigmoid(Y, crm):                # transform.py:141-153 ; (F,T,2)
    M = cRM_sigmoid_recover(crm)
    S = np.zeros(np.shape(M))
    S[..., 0] = M[..., 0] * Y[..., 0] - M[..., 1] * Y[..., 1]
    S[..., 1] = M[..., 0] * Y[..., 1] + M[..., 1] * Y[..., 0]
    return S


def batch_fast_icRM_sigmoid(Y, crm, a=0.1, b=0):
    """transform.py:156-169 on torch tensors (B,2,F,T); differentiable."""
    import torch
    M = 1.0 / a * (torch.log(crm / (1 - crm + 1e-8) + 1e-10) + b)
    r = M[:, 0] * Y[:, 0] - M[:, 1] * Y[:, 1]
    i = M[:, 0] * Y[:, 1] + M[:, 1] * Y[:, 0]
    return torch.stack([r, i], dim=1)


def stft_batch(waves):
    """(B,L) float32 -> (B,2,256,T) float32, the layout Dataset items use
    (M2/dataset.py:255-259: transpose((2,0,1)) of fast_stft)."""
    out = [fast_stft(w).transpose(2, 0, 1) for w in np.asarray(waves)]
    return np.stack(out).astype(np.float32)


def istft_batch(spec):
    """(B,2,256,T) -> (B,158*(T-1)) float32."""
    spec = np.asarray(spec)
    return np.stack([fast_istft(s.transpose(1, 2, 0)) for s in spec]).astype(np.float32)
