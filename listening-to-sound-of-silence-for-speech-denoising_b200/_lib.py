"""ctypes binding of libsos_b200.so (the C ABI declared in include/sos_b200.h).

There is no CPU or PyTorch fallback: if the library is missing, `lib()` raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsos_b200.so")
CSRC = os.path.join(_HERE, "csrc")

c_f = C.c_void_p          # device pointers travel as integers
i64 = C.c_int64
i32p = C.c_void_p


class ConvArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("wk", C.c_void_p), ("y", C.c_void_p),
                ("tap_dh", C.POINTER(C.c_int32)), ("tap_dw", C.POINTER(C.c_int32)),
                ("N", i64), ("H", i64), ("W", i64), ("Cin", i64),
                ("Cout", i64), ("OH", i64), ("OW", i64),
                ("ntaps", i64), ("stride", i64),
                ("YH", i64), ("YW", i64), ("Cy", i64), ("y_coff", i64),
                ("osh", i64), ("osw", i64), ("oph", i64), ("opw", i64),
                ("epi_scale", C.c_void_p), ("epi_shift", C.c_void_p),
                ("act", i64), ("slope", C.c_void_p),
                ("force_plan", i64), ("plan_out", C.POINTER(C.c_int32)),
                ("stats_partial", C.c_void_p), ("stats_channels", i64), ("stats_rows_out", C.POINTER(C.c_int32)),
                ("x_dtype", i64), ("y_dtype", i64), ("out_scale", C.c_void_p),
                ("bnr_y", C.c_void_p), ("bnr_scale", C.c_void_p), ("bnr_shift", C.c_void_p), ("bnr_mean", C.c_void_p), ("bnr_invstd", C.c_void_p),
                ("bnr_partial", C.c_void_p), ("bnr_channels", i64), ("bnr_rows_out", C.POINTER(C.c_int32))]


class PackDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out_half", C.c_void_p), ("rows", i64), ("K", i64), ("KP", i64), ("row_stride", i64), ("k_stride", i64),
                ("ntaps", i64), ("tap_off", C.c_int32 * 50)]


class WgradArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p),
                ("tap_dh", C.POINTER(C.c_int32)), ("tap_dw", C.POINTER(C.c_int32)),
                ("N", i64), ("H", i64), ("W", i64), ("Cin", i64),
                ("Cout", i64), ("OH", i64), ("OW", i64), ("Cdy", i64), ("dy_coff", i64),
                ("ntaps", i64), ("stride", i64),
                ("force_plan", i64), ("plan_out", C.POINTER(C.c_int32)),
                ("dtype", i64), ("out_scale", C.c_void_p)]


S = C.c_void_p  # cudaStream_t
_SIGS = {
    "sos_last_error": (C.c_char_p, []),
    "sos_version": (C.c_int, []),
    "sos_init": (C.c_int, []),
    "sos_stft_forward": (C.c_int, [c_f, i64, i64, c_f, c_f, i64, c_f, C.c_double, C.c_int, S]),
    "sos_istft_forward": (C.c_int, [c_f, c_f, i64, i64, c_f, c_f, S]),
    "sos_gate_wave": (C.c_int, [c_f, i64, i64, c_f, i64, c_f, C.c_double, C.c_int, c_f, c_f, S]),
    "sos_icrm_forward": (C.c_int, [c_f, c_f, c_f, i64, i64, C.c_float, C.c_float, S]),
    "sos_icrm_backward": (C.c_int, [c_f, c_f, c_f, c_f, i64, i64, C.c_float, S]),
    "sos_add_signals": (C.c_int, [c_f, c_f, c_f, i64, i64, C.c_float, c_f, c_f, c_f, S]),
    "sos_crm_forward": (C.c_int, [c_f, c_f, c_f, i64, i64, C.c_float, C.c_float, S]),
    "sos_resample": (C.c_int, [c_f, i64, i64, c_f, i64, c_f, c_f, i64, i64, C.c_double, c_f, S]),
    "sos_metric_frames": (C.c_int, [i64, i64]),
    "sos_wss": (C.c_int, [c_f, c_f, i64, i64, i64, C.c_double, c_f, S]),
    "sos_llr": (C.c_int, [c_f, c_f, i64, i64, i64, c_f, S]),
    "sos_ssnr": (C.c_int, [c_f, c_f, i64, i64, i64, C.c_double, C.c_float, C.c_float, C.c_double, C.c_int, c_f, c_f, S]),
    "sos_mse_fwd_bwd": (C.c_int, [c_f, c_f, i64, c_f, c_f, C.c_float, S]),
    "sos_bce_logits_fwd_bwd": (C.c_int, [c_f, c_f, i64, c_f, c_f, C.c_float, S]),
    "sos_round_tf32": (C.c_int, [c_f, i64, S]),
    "sos_adam_step": (C.c_int, [c_f, c_f, c_f, c_f, i64, C.c_float, C.c_float, C.c_float, C.c_float, i64, C.c_float, S]),
    "sos_adam_step_dev": (C.c_int, [c_f, c_f, c_f, c_f, i64, c_f, C.c_float, C.c_float, C.c_float, C.c_float, S]),
    "sos_plan_cache_stats": (None, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sos_bn_partial_blocks": (C.c_int, [i64, i64]),
    "sos_bn_stats": (C.c_int, [c_f, i64, i64, c_f, S]),
    "sos_bn_finalize": (C.c_int, [c_f, i64, i64, c_f, c_f, C.c_float, C.c_float, c_f, c_f, c_f, c_f, c_f, c_f, S]),
    "sos_bn_finalize_partial": (C.c_int, [c_f, i64, i64, i64, c_f, c_f, C.c_float, C.c_float, c_f, c_f, c_f, c_f, c_f, c_f, S]),
    "sos_conv_stats_rows": (C.c_int, []),
    "sos_bn_eval_coeffs": (C.c_int, [i64, c_f, c_f, c_f, c_f, C.c_float, c_f, c_f, S]),
    "sos_bn_act": (C.c_int, [c_f, c_f, i32p, i64, i64, c_f, c_f, C.c_int, c_f, S]),
    "sos_bn_act_backward": (C.c_int, [c_f, i32p, c_f, c_f, i64, i64, c_f, c_f, c_f, c_f, C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_f, S]),
    "sos_bn_act_half": (C.c_int, [c_f, C.c_int, c_f, i64, i64, c_f, c_f, C.c_int, c_f, S]),
    "sos_bn_act_backward_half": (C.c_int, [c_f, C.c_int, c_f, c_f, C.c_int, c_f, i64, i64, c_f, c_f, c_f, c_f, C.c_int, c_f, c_f, c_f, c_f, c_f,
                                           c_f, c_f, c_f, C.c_int, i64, S]),
    "sos_bn_act_backward_half_pre": (C.c_int, [c_f, C.c_int, c_f, c_f, C.c_int, c_f, i64, i64, c_f, c_f, c_f, c_f, C.c_int, c_f, c_f, i64, c_f, c_f, c_f,
                                               c_f, c_f, c_f, C.c_int, i64, S]),
    "sos_nchw_to_nhwc_half": (C.c_int, [c_f, i64, i64, i64, i64, c_f, i64, S]),
    "sos_accumulate_wgrad": (C.c_int, [c_f, i64, i64, i64, i64, i64, c_f, S]),
    "sos_accumulate_wgrad_clear": (C.c_int, [c_f, i64, i64, i64, i64, i64, c_f, S]),
    "sos_to_half": (C.c_int, [c_f, i64, i64, c_f, i64, c_f, S]),
    "sos_affine_act_backward": (C.c_int, [c_f, i32p, c_f, c_f, i64, i64, c_f, c_f, C.c_int, c_f, S]),
    "sos_nchw_to_nhwc": (C.c_int, [c_f, i64, i64, c_f, i32p, i64, S]),
    "sos_nhwc_to_nchw": (C.c_int, [c_f, i32p, i64, i64, c_f, S]),
    "sos_copy_view": (C.c_int, [c_f, i32p, c_f, i32p, i64, i64, C.c_int, S]),
    "sos_copy_view_backward": (C.c_int, [c_f, i32p, c_f, i32p, i64, i64, S]),
    "sos_im2col_half": (C.c_int, [c_f, i64, i64, i64, i64, i64, i32p, i32p, i64, i64, c_f, i64, S]),
    "sos_copy_view_fold": (C.c_int, [c_f, i32p, c_f, i32p, i64, i64, S]),
    "sos_reflect_fill": (C.c_int, [c_f, i64, i64, i64, i64, i64, S]),
    "sos_reflect_fold": (C.c_int, [c_f, i64, i64, i64, i64, i64, S]),
    "sos_feat_to_seq": (C.c_int, [c_f, i64, i64, i64, i64, c_f, i64, i64, i64, S]),
    "sos_feat_to_seq_backward": (C.c_int, [c_f, i64, i64, i64, i64, c_f, i64, i64, i64, S]),
    "sos_transpose": (C.c_int, [c_f, i64, i64, c_f, S]),
    "sos_split_tf32": (C.c_int, [c_f, i64, i64, i64, i64, i64, i64, c_f, i64, i64, C.c_int, i64, i64, i64, S]),
    "sos_istft_ola": (C.c_int, [c_f, i64, i64, c_f, S]),
    "sos_axpy": (C.c_int, [c_f, c_f, i64, C.c_float, c_f, S]),
    "sos_seq_map": (C.c_int, [c_f, c_f, i64, i64, i64, i64, C.c_int, S]),
    "sos_bias_act": (C.c_int, [c_f, i64, i64, i64, c_f, C.c_int, S]),
    "sos_bias_act_backward": (C.c_int, [c_f, c_f, c_f, i64, i64, i64, C.c_int, c_f, S]),
    "sos_pack_conv_weight": (C.c_int, [c_f, i64, i64, i64, i64, i64, i64, C.c_int, c_f, S]),
    "sos_pack_taps": (C.c_int, [c_f, i64, i64, i64, i64, i64, i64, C.POINTER(C.c_int32), C.c_int, c_f, S]),
    "sos_pack_taps_half": (C.c_int, [c_f, i64, i64, i64, i64, i64, i64, C.POINTER(C.c_int32), c_f, S]),
    "sos_pack_taps_half_multi": (C.c_int, [c_f, i64, S]),
    "sos_unpack_wgrad": (C.c_int, [c_f, i64, i64, i64, i64, c_f, C.c_int, S]),
    "sos_conv2d_tc": (C.c_int, [C.POINTER(ConvArgs), S]),
    "sos_conv2d_plan": (C.c_int, [C.POINTER(ConvArgs), C.POINTER(C.c_int32)]),
    "sos_conv2d_wgrad": (C.c_int, [C.POINTER(WgradArgs), S]),
    "sos_lstm_forward": (C.c_int, [c_f, c_f, i64, i64, i64, c_f, c_f, c_f, S]),
    "sos_lstm_backward": (C.c_int, [c_f, c_f, c_f, c_f, c_f, i64, i64, i64, c_f, c_f, c_f, S]),
}
EXPORTS = tuple(_SIGS)

_lib = None


class SosError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libsos_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise SosError("building libsos_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SosError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the sos_b200 hot path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().sos_last_error()
        raise SosError(f"{what or 'libsos_b200'} failed ({code}): {msg.decode() if msg else ''}")


launch_count = 0       # kernels launched through this binding (bench.py's gpu_launches)
