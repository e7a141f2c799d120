"""The level model of the 16-bit epilogue (csrc/estimate.h, estimate.cpp): a piecewise-quadratic approximation of
"linear sample -> output code" that settles almost every sample without the 65 536-entry table.  Whatever it settles
must equal the reference's rule (src/imagew-main.c:669-765 without dither: nearest code in LINEAR space, ties to the
upper code, values above 0.999999999 to the top code, clamped to [0, 1] first); everything else must be left to the
table (-1).  Host code, the kernel's own arithmetic (the hook runs the same inline function)."""
import ctypes as C

import numpy as np
import pytest

from imageworsener_b200 import libimagew as L


def _table(lib, cstype, top):
    cs = L.iw_csdescr()
    cs.cstype = cstype
    return np.array([lib.dll.iw_convert_sample_to_linear(k / top, C.byref(cs)) for k in range(top + 1)])


def _exact(thr, v):
    top = len(thr) - 1
    s = np.clip(v, 0.0, 1.0)
    lo = np.searchsorted(thr, s, side="right") - 1
    lo = np.clip(lo, 0, top)
    nxt = np.minimum(lo + 1, top)
    up = (thr[nxt] - s) <= (s - thr[lo])
    code = np.where(lo >= top, top, lo + up.astype(np.int64))
    return np.where(s > 0.999999999, top, code)


def _model(lib, cstype, thr, v):
    fn = lib.dll.iwb200_debug_level_model
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    code = np.empty(len(v), dtype=np.int32)
    stats = np.zeros(3)
    v = np.ascontiguousarray(v, dtype=np.float64)
    ok = fn(cstype, thr.ctypes.data, len(thr), v.ctypes.data, len(v), code.ctypes.data, stats.ctypes.data)
    return ok, code, stats


def _nudge(x, ulps):
    return (x.view(np.int64) + ulps).view(np.float64)


@pytest.mark.parametrize("cstype", [L.IW_CSTYPE_SRGB, L.IW_CSTYPE_REC709], ids=["srgb", "rec709"])
@pytest.mark.parametrize("top", [65535, 4095, 1023])
def test_whatever_the_model_settles_is_the_reference_rule(emu, cstype, top):
    thr = _table(emu, cstype, top)
    rng = np.random.default_rng(top + cstype)
    mid = (thr[:-1] + thr[1:]) * 0.5
    q = rng.integers(0, top, 200_000)
    samples = np.concatenate([
        rng.random(400_000),                                   # everyday samples
        rng.random(100_000) * 0.01, rng.random(100_000) * 3e-4,   # the toe, and below the first octave of the model
        _nudge(mid[q], rng.integers(-2000, 2001, len(q))),     # razor-edge: within 2000 ulps of a tie
        _nudge(thr[q + 1], rng.integers(-3, 4, len(q))),       # at the table's own entries
        -rng.random(1000), 1.0 + rng.random(1000), np.array([0.0, 1.0, 5e-324, 1e-300, 0.999999999, 0.9999999995]),
    ])
    ok, code, stats = _model(emu, cstype, thr, samples)
    assert ok == 1
    want = _exact(thr, samples)
    settled = code >= 0
    assert np.array_equal(code[settled], want[settled])
    # worth having: nearly all everyday samples are settled; the ties are not
    assert settled[:400_000].mean() > (0.99 if top == 65535 else 0.98)
    assert not settled[600_000:800_000].any()
    assert stats[0] < 5e-3 and stats[2] < 50


def test_model_leaves_nan_and_infinities_to_the_table(emu):
    thr = _table(emu, L.IW_CSTYPE_SRGB, 65535)
    ok, code, _ = _model(emu, L.IW_CSTYPE_SRGB, thr, np.array([np.nan, np.inf, -np.inf, -0.0]))
    assert ok == 1
    assert code[0] == -1 and code[1] == -1
    assert code[2] in (-1, 0) and code[3] == 0


def test_no_model_for_power_law_or_tiny_tables(emu):
    thr = _table(emu, L.IW_CSTYPE_SRGB, 65535)
    assert _model(emu, L.IW_CSTYPE_GAMMA, thr, np.array([0.5]))[0] == 0
    assert _model(emu, L.IW_CSTYPE_SRGB, _table(emu, L.IW_CSTYPE_SRGB, 63), np.array([0.5]))[0] == 0
