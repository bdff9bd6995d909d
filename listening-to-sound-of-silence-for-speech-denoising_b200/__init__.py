"""sos_b200 -- B200-native (sm_100a) hot path of "Listening to Sounds of Silence for Speech Denoising".

Host code mirrors the reference's Python surface (get_network / forward / transform.* / agent) and calls
hand-written CUDA through the C ABI of libsos_b200.so (include/sos_b200.h).  No CPU fallback.
"""
from . import _lib                                     # noqa: F401
from ._lib import SosError, build, lib                 # noqa: F401

__all__ = ["SosError", "build", "lib"]
