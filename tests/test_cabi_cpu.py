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
