"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/sos_b200.h declares (no compute calls:
there is no GPU here).  Also: the product path refuses to run without a CUDA device instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import sos_b200
    sos_b200.build()
    return sos_b200


def _declared():
    src = open(os.path.join(ROOT, "include", "sos_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sos_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built):
    lib = ctypes.CDLL(built._lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/sos_b200.h but not exported by libsos_b200.so"


def test_binding_covers_header(built):
    assert sorted(built._lib.EXPORTS) == _declared()
    built.lib()                                    # resolves every symbol with its ctypes signature
    assert built.lib().sos_version() >= 1


def test_argument_errors_are_reported(built):
    """Argument validation happens on the host before any CUDA call: error code + message, no crash."""
    lib = built.lib()
    assert lib.sos_bn_partial_blocks(1000, 6) == 0                       # channels must be a multiple of 4
    assert lib.sos_bn_partial_blocks(1 << 20, 96) > 0
    rc = lib.sos_stft_forward(None, 1, 32000, None, None, 0, None, 0.0, 0, None)
    assert rc == -1 and b"sos_stft_forward" in lib.sos_last_error()
    rc = lib.sos_conv2d_tc(None, None)
    assert rc == -1 and b"null" in lib.sos_last_error()


def test_argument_errors_of_the_newer_entry_points(built):
    """Host-side validation of the half-operand producers, the item builder and the metrics: bad arguments give -1 and a
    message naming the entry point, before any CUDA call."""
    lib = built.lib()
    one = ctypes.c_void_p(16)                                            # a non-null dummy pointer (never dereferenced on these paths)
    assert lib.sos_to_half(one, 10, 8, one, 12, None, None) == -1 and b"sos_to_half" in lib.sos_last_error()          # cd % 8
    assert lib.sos_to_half(one, 10, 16, one, 8, None, None) == -1                                                      # cd < cs
    assert lib.sos_bn_act_half(one, 0, one, 10, 12, one, one, 1, None, None) == -1 and b"channels % 8" in lib.sos_last_error()
    assert lib.sos_bn_act_half(one, 1, one, 10, 16, one, one, 2, None, None) == -1 and b"PReLU" in lib.sos_last_error()  # slope missing
    assert lib.sos_bn_act_half(one, 5, one, 10, 16, one, one, 1, None, None) == -1 and b"unknown y type" in lib.sos_last_error()
    assert lib.sos_bn_act_backward_half(one, 3, None, one, 0, one, 100, 16, one, one, one, one, 1, None, one, one, one, None, one, one, one, 0, 0,
                                        None) == -1 and b"unknown dz" in lib.sos_last_error()
    assert lib.sos_adam_step_dev(one, one, one, one, 10, one, 0.9, 0.999, 1e-8, 1.0, None) == -1 and b"multiple of 4" in lib.sos_last_error()
    assert lib.sos_add_signals(one, one, None, 4, 100, 0.5, one, one, one, None) == -1 and b"sos_add_signals" in lib.sos_last_error()
    assert lib.sos_add_signals(one, one, one, 0, 100, 0.5, one, one, one, None) == -1                                  # empty batch
    assert lib.sos_crm_forward(one, one, one, 2, 0, 0.1, 0.0, None) == -1 and b"sos_crm_forward" in lib.sos_last_error()
    assert lib.sos_ssnr(one, one, 1, 100, 16000, 0.0, -10.0, 35.0, 1e-10, 0, one, one, None) == -1 and b"sos_ssnr" in lib.sos_last_error()
    assert lib.sos_ssnr(one, one, 1, 100, 100, 30.0, -10.0, 35.0, 1e-10, 0, one, one, None) == -1 and b"too short" in lib.sos_last_error()
    assert lib.sos_nchw_to_nhwc_half(one, 1, 2, 4, 4, one, 12, None) == -1 and b"multiple of 8" in lib.sos_last_error()
    assert lib.sos_accumulate_wgrad(one, 9, 8, 8, 16, 8, one, None) == -1                                              # rows > rows_padded
    assert lib.sos_pack_taps_half(one, 4, 4, 8, 36, 9, 50, (ctypes.c_int32 * 50)(), one, None) == -1                   # > 49 taps
    a = built._lib.ConvArgs()
    a.x, a.wk, a.y = 16, 16, 16
    dh = (ctypes.c_int32 * 1)(0)
    a.tap_dh, a.tap_dw = dh, dh
    a.N, a.H, a.W, a.Cin, a.Cout, a.OH, a.OW, a.ntaps, a.stride = 1, 8, 8, 12, 16, 8, 8, 1, 1
    a.YH, a.YW, a.Cy, a.osh, a.osw = 8, 8, 16, 1, 1
    a.force_plan = -1
    assert lib.sos_conv2d_tc(ctypes.byref(a), None) == -1 and b"multiple of 8" in lib.sos_last_error()                # Cin = 12
    a.Cin, a.x_dtype = 16, 7
    assert lib.sos_conv2d_tc(ctypes.byref(a), None) == -1 and b"unknown operand" in lib.sos_last_error()
    g = built._lib.WgradArgs()
    g.x, g.dy, g.dw = 16, 16, 16
    g.tap_dh, g.tap_dw = dh, dh
    g.N, g.H, g.W, g.Cin, g.Cout, g.OH, g.OW, g.Cdy, g.ntaps, g.stride, g.dtype = 1, 8, 8, 16, 16, 8, 8, 12, 1, 1, 1
    g.force_plan = -1
    assert lib.sos_conv2d_wgrad(ctypes.byref(g), None) == -1 and b"dy channel slice" in lib.sos_last_error()           # half: Cdy % 8


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from sos_b200 import ops, transform
    with pytest.raises(built.SosError):
        ops.init()
    with pytest.raises(built.SosError):
        transform.stft_batch(torch.zeros(1, 32000))


def test_sass_is_blackwell_native(built):
    """The shipped library contains tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG / UTMASTG) -- no legacy HMMA."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", built._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic
    assert " HMMA" not in sass and "HGMMA" not in sass
