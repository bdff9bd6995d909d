"""CPU: the host-side planner of the convolution entry point (sos_conv2d_plan: pure shape logic, no CUDA call) routes the layers of
the two networks the way DESIGN.md 4.1 / 4.1b describe: the 48-channel dilated 5x5 layers (half in / half out, H-dilation <= 16) to the
row-streaming kernel, the wide layers to CTA pairs with five pipeline stages, the epilogue-bound small layers to single CTAs, and
48-channel inputs the tap GEMM still serves as ONE zero-tailed 64-channel chunk."""
import ctypes as C

import pytest


@pytest.fixture(scope="module")
def lib():
    import sos_b200
    sos_b200.build()
    return sos_b200.lib(), sos_b200._lib


def _plan(lib, mod, Cin, Cout, k, d, H=256, W=203, N=32, y_half=1, valid=False, negate=False):
    offs = [((a - (0 if valid else (k[0] - 1) // 2)) * d[0], (b - (0 if valid else (k[1] - 1) // 2)) * d[1]) for a in range(k[0]) for b in range(k[1])]
    if negate:                                   # the tap list of a data gradient: same weight taps, mirrored offsets
        offs = [(-a, -b) for a, b in offs]
    OH, OW = (H - (k[0] - 1) * d[0], W - (k[1] - 1) * d[1]) if valid else (H, W)
    a = mod.ConvArgs()
    dh = (C.c_int32 * len(offs))(*[o[0] for o in offs])
    dw = (C.c_int32 * len(offs))(*[o[1] for o in offs])
    a.tap_dh, a.tap_dw = dh, dw
    a.N, a.H, a.W, a.Cin, a.Cout, a.OH, a.OW, a.ntaps, a.stride = N, H, W, Cin, Cout, OH, OW, len(offs), 1
    a.YH, a.YW, a.Cy, a.osh, a.osw = OH, OW, (Cout + 7) // 8 * 8, 1, 1
    a.force_plan, a.x_dtype, a.y_dtype = -1, 1, y_half
    info = (C.c_int32 * 16)()
    assert lib.sos_conv2d_plan(C.byref(a), info) == 0, lib.sos_last_error()
    names = "kernel share g S groups stages stage_bytes grid cbe chunks N ec FB SB n_stg smem".split()
    return dict(zip(names, info))


@pytest.mark.parametrize("d", [(1, 1), (2, 1), (4, 1), (8, 1), (16, 1), (2, 2), (4, 4), (8, 8), (16, 16)])
def test_narrow_dilated_layers_go_to_the_row_kernel(lib, d):
    assert _plan(*lib, 48, 48, (5, 5), d)["kernel"] == 2


@pytest.mark.parametrize("d", [(32, 1), (32, 32)])
def test_h_dilation_32_stays_on_the_tap_gemm_on_pairs(lib, d):
    p = _plan(*lib, 48, 48, (5, 5), d)                 # a 256-row box does not fit next to the resident weights
    assert p["kernel"] in (0, 1) and p["cbe"] == 64 and p["chunks"] == 1 and p["n_stg"] >= 100


def test_row_kernel_needs_half_output_and_whole_h_blocks(lib):
    assert _plan(*lib, 48, 48, (5, 5), (1, 1), y_half=0)["kernel"] != 2
    assert _plan(*lib, 48, 48, (5, 5), (1, 1), H=192)["kernel"] != 2
    assert _plan(*lib, 48, 48, (7, 1), (1, 1))["kernel"] != 2            # no taps along W to stack
    assert _plan(*lib, 96, 96, (5, 5), (1, 1))["kernel"] != 2            # N = 5 * 96 > 256


@pytest.mark.parametrize("Cin,Cout,k,d,H,W,valid", [(96, 96, (5, 5), (1, 1), 256, 203, False), (96, 96, (5, 5), (8, 8), 256, 203, False),
                                                     (128, 128, (5, 5), (1, 1), 132, 106, True), (256, 256, (3, 3), (1, 1), 66, 53, True)])
def test_wide_layers_run_on_cta_pairs(lib, Cin, Cout, k, d, H, W, valid):
    p = _plan(*lib, Cin, Cout, k, d, H=H, W=W, valid=valid)
    assert p["n_stg"] >= 100, p                         # +100: cta_group::2
    assert p["stages"] >= 4 and p["smem"] <= 227 * 1024


@pytest.mark.parametrize("Cin,Cout,k", [(16, 96, (1, 7)), (96, 8, (1, 1)), (96, 96, (7, 1))])
def test_epilogue_bound_layers_stay_on_single_ctas(lib, Cin, Cout, k):
    p = _plan(*lib, Cin, Cout, k, (1, 1))
    assert p["n_stg"] < 100 and p["stages"] >= 2
    if Cin == 16:
        assert p["cbe"] == 16                           # 16 real channels: no half-empty 32-wide chunk (TMA zero fill is not free)


def test_store_bound_layers_stage_wide_rows_and_heavy_layers_merge_weight_loads(lib):
    """info[1] bits: +2 = half outputs staged as 128-byte rows of 64 channels (store-bound layers only: the tensor-bound ones keep
    their shared memory for operand stages), +4 = the weight tiles of a stage's sub-taps arrive with ONE TMA load (equally spaced
    taps: the k x k layers; a 1 x 1 layer has nothing to merge)."""
    heavy = _plan(*lib, 96, 96, (5, 5), (1, 1))
    assert heavy["share"] & 4 and not heavy["share"] & 2, heavy
    light = _plan(*lib, 16, 96, (1, 1), (1, 1))                      # 96 -> 8 data gradient: K = 16, N = 96, one tap
    assert light["share"] & 2 and not light["share"] & 4, light
    mid = _plan(*lib, 256, 256, (3, 3), (1, 1), 66, 53, valid=True)
    assert mid["share"] & 4, mid
    dgrad = _plan(*lib, 96, 96, (5, 5), (2, 2), negate=True)                 # taps of a box run DOWN the packed rows: loaded from the last one
    assert dgrad["share"] & 4, dgrad
