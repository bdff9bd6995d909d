"""CPU restatement of the reference's audio loading: `librosa.load(path, sr=14000)` (M2/predict.py:303, M1/dataset.py:226,
M1/create_data_from_pred.py:130) = read -> float32 in [-1, 1) -> mono (mean of the channels) -> `resampy.resample(y, sr_orig,
sr_new, filter='kaiser_best')` -> fix_length(ceil(n * ratio)).  TEST INFRASTRUCTURE.

Third-party code absent from /root/reference (and not installable offline): librosa==0.7.1 (requirements.txt:4) and its
dependency resampy (0.2.x).  Their published algorithm is restated here:

* filter (resampy/filters.py `sinc_window`, data file kaiser_best.npz): num_zeros = 64, precision = 9 (512 table samples per zero
  crossing), rolloff = 0.9475937167399596, Kaiser window beta = 14.769656459379492;
      interp_win[i] = kaiser(2 n + 1, beta)[n + i] * rolloff * sinc(rolloff * i / 512),  i = 0 .. n = 64 * 512
  scaled by the sample ratio when down-sampling; interp_delta = forward differences (last entry 0);
* interpolation (resampy/interpn.py `resample_f`, J. O. Smith's band-limited interpolation): for output sample t at input time
  t / ratio = n + frac: left wing over x[n], x[n-1], ...; right wing over x[n+1], x[n+2], ... with linearly interpolated filter
  values; the accumulator has the dtype of x (float32 after librosa.load), every `+=` rounds to it; the time register is advanced by
  repeated addition of 1 / ratio.

PARITY UNPINNED for this leg: there is no resampy binary here to compare with; the GPU kernel is checked against this restatement,
and this restatement against scipy.signal.resample_poly (a different low-pass design: agreement to ~1e-3 in band)."""
import numpy as np

NUM_ZEROS, PRECISION = 64, 9
ROLLOFF, BETA = 0.9475937167399596, 14.769656459379492


def kaiser_best_filter():
    from scipy.signal.windows import kaiser
    num_bits = 2 ** PRECISION
    n = num_bits * NUM_ZEROS
    sinc_win = ROLLOFF * np.sinc(ROLLOFF * np.linspace(0, NUM_ZEROS, num=n + 1, endpoint=True))
    taper = kaiser(2 * n + 1, BETA)[n:]
    return taper * sinc_win, num_bits


def resample(x, sr_orig, sr_new):
    """resampy.resample(x, sr_orig, sr_new, filter='kaiser_best') for a 1-D float32 signal (loops over the output samples like
    resample_f; vectorised over the filter taps)."""
    x = np.asarray(x)
    ratio = float(sr_new) / sr_orig
    n_out = int(x.shape[0] * ratio)
    interp_win, num_table = kaiser_best_filter()
    if ratio < 1:
        interp_win = interp_win * ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    scale = min(1.0, ratio)
    time_increment = 1.0 / ratio
    index_step = int(scale * num_table)
    nwin, n_orig = interp_win.shape[0], x.shape[0]
    y = np.zeros(n_out, dtype=x.dtype)
    time_register = 0.0
    xd = x.astype(np.float64)
    for t in range(n_out):
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i_max = min(n + 1, (nwin - offset) // index_step)
        acc = y.dtype.type(0)
        if i_max > 0:
            idx = offset + np.arange(i_max) * index_step
            terms = (interp_win[idx] + eta * interp_delta[idx]) * xd[n - np.arange(i_max)]
            for v in terms:                               # `y[t] += weight * x[n - i]`: one rounding to x's dtype per tap
                acc = y.dtype.type(np.float64(acc) + v)
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k_max = min(n_orig - n - 1, (nwin - offset) // index_step)
        if k_max > 0:
            idx = offset + np.arange(k_max) * index_step
            terms = (interp_win[idx] + eta * interp_delta[idx]) * xd[n + 1 + np.arange(k_max)]
            for v in terms:
                acc = y.dtype.type(np.float64(acc) + v)
        y[t] = acc
        time_register += time_increment
    return y


def librosa_load_array(x, sr_file, sr=14000):
    """librosa.load on already-decoded samples: x (n,) or (n, channels) int16 / float -> (float32 mono at `sr`, sr)."""
    x = np.asarray(x)
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(np.iinfo(x.dtype).max + 1)
    x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)                                # to_mono
    if sr is None or sr == sr_file:
        return x, sr_file
    ratio = float(sr) / sr_file
    n_samples = int(np.ceil(x.shape[-1] * ratio))
    y = resample(x, sr_file, sr)
    if y.shape[0] < n_samples:                            # util.fix_length
        y = np.pad(y, (0, n_samples - y.shape[0]))
    return np.ascontiguousarray(y[:n_samples], dtype=np.float32), sr
