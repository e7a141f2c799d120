"""Host-side planning of the streaming kernels' launches (csrc/stream2_host.cuh: iwb2_plan_parts), without a GPU.

A launch over outputs [lo, hi) is split into `n_u` uniform parts covering [u_lo, u_hi) and general parts covering
the rest.  Whatever the schedule looks like, the parts must tile [lo, hi) exactly once, respect the alignment, and
a range may only be called uniform where the window ends really advance by the constant integer step.
"""
import ctypes as C
import os

import numpy as np
import pytest

from imageworsener_b200 import libimagew as L


@pytest.fixture(scope="module")
def plan():
    dll = C.CDLL(L.PRODUCT_SO, mode=os.RTLD_LOCAL)
    fn = dll.iwb200_debug_plan_parts
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]

    def call(last, lo, hi, base, align, S, want, override, g_min, step_wanted, first=None):
        last = np.ascontiguousarray(last, np.int32)
        if first is None:                                          # windows of S*step - 1 inputs: slots are always free in time
            first = last - (S * max(step_wanted, 1) - 2)
        first = np.ascontiguousarray(np.maximum(first, 0), np.int32)
        out = np.zeros(9, np.int32)
        assert fn(first.ctypes.data, last.ctypes.data, len(last), lo, hi, base, align, S, want, override, g_min, step_wanted, out.ctypes.data)
        return dict(zip("step n_parts n_u u_lo u_hi u_len g_len u_ref_o u_ref_last".split(), map(int, out)))
    return call


def _parts(p, lo, hi):
    """Decode the parts exactly as the kernels do (stream2_dev.cuh: s2_decode_part)."""
    out = []
    for i in range(p["n_parts"]):
        if i < p["n_u"]:
            a = p["u_lo"] + i * p["u_len"]
            out.append((a, min(a + p["u_len"], p["u_hi"]), True))
            continue
        g = i - p["n_u"]
        nh = (p["u_lo"] - lo + p["g_len"] - 1) // p["g_len"]
        if g < nh:
            a = lo + g * p["g_len"]
            out.append((a, min(a + p["g_len"], p["u_lo"]), False))
        else:
            a = p["u_hi"] + (g - nh) * p["g_len"]
            out.append((a, min(a + p["g_len"], hi), False))
    return out


def _sched(n_out, step, jitter_head, jitter_tail, rng):
    """Window ends: a constant integer step in the middle, irregular (but monotone) near both borders."""
    last = np.arange(n_out) * step + 5
    for k in range(jitter_head):
        last[k] = max(0, last[k] - rng.integers(0, 2))
    for k in range(n_out - jitter_tail, n_out):
        last[k] = last[n_out - jitter_tail - 1] + (k - (n_out - jitter_tail - 1)) * max(1, step - 1)
    return np.maximum.accumulate(last)


@pytest.mark.parametrize("seed", range(40))
def test_parts_tile_the_launch_and_uniform_ranges_are_uniform(plan, seed):
    rng = np.random.default_rng(seed)
    S = int(rng.choice([2, 4, 6, 8]))
    step = int(rng.choice([1, 2, 3, 4]))
    n_out = int(rng.integers(40, 3000))
    last = _sched(n_out, step, int(rng.integers(0, 6)), int(rng.integers(0, 6)), rng)
    lo = int(rng.integers(0, n_out // 3))
    hi = int(rng.integers(lo + 8, n_out + 1))
    align = int(rng.choice([1, 4, 8]))
    base = lo
    want = int(rng.integers(1, 60))
    p = plan(last, lo, hi, base, align, S, want, 0, 16, step)
    parts = _parts(p, lo, hi)
    # exact tiling, in order within each class, no empty part
    cover = np.zeros(n_out, np.int32)
    for a, b, uni in parts:
        assert lo <= a < b <= hi
        cover[a:b] += 1
        assert (a - base) % align == 0
    assert np.all(cover[lo:hi] == 1) and cover.sum() == hi - lo
    if p["step"]:
        assert p["step"] == step
        # the progression the kernels assume holds for every output of every uniform part
        for a, b, uni in parts:
            if uni:
                o = np.arange(a, b)
                assert np.array_equal(last[o], p["u_ref_last"] + p["step"] * (o - p["u_ref_o"]))
        assert all(b - a >= 1 for a, b, uni in parts if uni)
    else:
        assert p["n_u"] == 0


def test_small_images_get_short_parts_but_not_shorter_than_their_halo(plan):
    last = np.arange(144) * 7 + 20                  # 1080 -> 144 rows, far more CTA slots than rows
    p = plan(last, 0, 144, 0, 1, 5, 37, 0, 16, 0)
    assert p["n_u"] == 0 and p["g_len"] == 10 and p["n_parts"] == 15      # 2*S outputs per part


def test_override_and_no_uniform_request(plan):
    last = np.arange(1000) * 2 + 6
    p = plan(last, 0, 1000, 0, 8, 6, 10, 200, 32, 2)
    assert p["step"] == 2 and p["u_len"] == 200
    q = plan(last, 0, 1000, 0, 8, 6, 10, 0, 32, 0)
    assert q["step"] == 0 and q["n_u"] == 0 and q["n_parts"] == (1000 + q["g_len"] - 1) // q["g_len"]


def test_first_outputs_after_a_break_in_the_progression_are_not_uniform(plan):
    """333 -> 166 pixels with Lanczos3 (factor 2.006): window ends advance by 2 except one 3 in the middle, windows are
    13 inputs long and 6 are open at once.  The kernels restart a slot S*step = 12 inputs before its output ends, so
    the outputs right after the break -- whose slots were last used by outputs from BEFORE the break -- must be left
    to the general path (this was a real bug: their first tap was lost)."""
    n_in, n_out, S, step = 333, 166, 6, 2
    o = np.arange(n_out)
    pos = (o + 0.5) * (n_in / n_out) - 0.5
    rad = 3.0 * n_in / n_out
    first = np.maximum(np.ceil(pos - rad), 0).astype(np.int32)
    last = np.minimum(np.floor(pos + rad), n_in - 1).astype(np.int32)
    brk = int(np.flatnonzero(np.diff(last)[40:120] != 2)[0]) + 41     # first output after the break
    p = plan(last, 0, n_out, 0, 4, S, 20, 0, 16, step, first=first)
    GU = S * step
    # here every other window is 13 inputs long (> GU), so whatever the planner calls uniform must avoid them all
    for a, b, uni in _parts(p, 0, n_out):
        if uni:
            for q in range(a, b):
                assert last[q] == p["u_ref_last"] + step * (q - p["u_ref_o"])
                assert first[q] > last[q] - GU and (q - S < 0 or last[q - S] <= last[q] - GU), (q, brk)
    assert all(not (uni and a <= brk < b) for a, b, uni in _parts(p, 0, n_out))
    # A break the other way (one window end advances by 1 instead of 2) with 12-input windows: the S outputs after
    # it reuse slots whose previous outputs end one input LATER than the progression says -- not uniform either;
    # everything further away from the break is.
    brk = 40                                                     # (away from the middle, where the planner anchors)
    last2 = 2 * o + 6
    last2[brk:] -= 1
    first2 = np.maximum(last2 - 11, 0).astype(np.int32)
    p = plan(last2, 0, n_out, 0, 4, S, 20, 0, 16, step, first=first2)
    assert p["step"] == 2 and p["n_u"] >= 1
    parts = _parts(p, 0, n_out)
    assert all(not (uni and a < brk + S and b > brk) for a, b, uni in parts)
    assert sum(b - a for a, b, uni in parts if uni) >= 60
    for a, b, uni in parts:
        if uni:
            for q in range(a, b):
                assert last2[q] == p["u_ref_last"] + step * (q - p["u_ref_o"])
                assert q - S < 0 or last2[q - S] <= last2[q] - GU


# ---- streaming schedules from tap lists (pipeline.cu:build_sched) ---------------------------------------------

def _schedule(dll, lists, n_in):
    """lists[o] = [(src, weight), ...] in list order."""
    fn = dll.iwb200_debug_build_schedule
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 4
    n_out = len(lists)
    tf = np.zeros(n_out + 1, np.int32)
    for o, l in enumerate(lists):
        tf[o + 1] = tf[o] + len(l)
    src = np.array([s for l in lists for s, _ in l], np.int32)
    wt = np.array([w for l in lists for _, w in l], np.float64)
    S = C.c_int()
    first, last = np.zeros(n_out, np.int32), np.zeros(n_out, np.int32)
    w = np.full(n_in * 16, np.nan)
    ok = fn(tf.ctypes.data, src.ctypes.data, wt.ctypes.data, n_in, n_out, C.byref(S), first.ctypes.data, last.ctypes.data, w.ctypes.data)
    if not ok:
        return None
    return S.value, first, last, w[: n_in * S.value].reshape(n_in, S.value)


@pytest.fixture(scope="module")
def dll():
    return L.load_product().dll


def _check_schedule(lists, n_in, res):
    """Every tap's weight sits at (its source, its output's slot); everything else the slot sees while the window is
    open is exactly +0.0; windows cover their taps and are monotone; no two open windows share a slot."""
    S, first, last, w = res
    expect = np.zeros_like(w)
    for o, l in enumerate(lists):
        assert first[o] <= min(s for s, _ in l) and last[o] >= max(s for s, _ in l)
        for s, wt in l:
            expect[s, o % S] = wt
    assert np.array_equal(w, expect) and not np.signbit(w[w == 0]).any()
    assert (np.diff(first) >= 0).all() and (np.diff(last) >= 0).all()
    for o in range(S, len(lists)):
        assert last[o - S] < first[o], "slot reused while its previous window is still open"


def test_schedule_of_contiguous_windows(dll):
    lists = [[(2 * o + k, 0.25) for k in range(4) if 0 <= 2 * o + k - 1 < 40] for o in range(20)]
    lists = [[(s - 1, w) for s, w in l] for l in lists]
    res = _schedule(dll, lists, 40)
    assert res is not None and res[0] == 2
    _check_schedule(lists, 40, res)


def test_schedule_accepts_holes_inside_a_window(dll):
    """3x reduction with a cubic filter: the tap at |x| = 1 has weight exactly 0 and is absent from the list."""
    lists = []
    for o in range(30):
        c = 3 * o + 1
        taps = [(s, 0.1 + 0.01 * k) for k, s in enumerate(range(c - 5, c + 6)) if 0 <= s < 90 and s not in (c - 3, c + 3)]
        lists.append(taps)
    res = _schedule(dll, lists, 90)
    assert res is not None
    _check_schedule(lists, 90, res)
    S, first, last, w = res
    assert w[3 * 10 + 1 + 3, 10 % S] == 0.0                  # the hole carries a zero weight


def test_schedule_widens_windows_whose_end_tap_was_dropped(dll):
    """Enlargement: an output whose first (last) tap was an exact zero starts after (ends before) its neighbour."""
    lists = [[(0, .5), (1, .5)], [(1, 1.0)], [(0, .1), (1, .8), (2, .1)], [(1, .5), (2, .5)], [(1, .2), (2, .6), (3, .2)],
             [(2, 1.0)], [(2, .5), (3, .5)], [(2, .1), (3, .9)]]
    res = _schedule(dll, lists, 4)
    assert res is not None
    _check_schedule(lists, 4, res)
    S, first, last, w = res
    assert first[1] == 0 and last[5] == 3                      # widened to keep starts and ends monotone


def test_schedule_rejects_what_the_streaming_kernels_cannot_do(dll):
    ok = [[(o, 0.5), (o + 1, 0.5)] for o in range(10)]
    assert _schedule(dll, ok, 11) is not None
    assert _schedule(dll, [[(-1, 0.5), (0, 0.5)]] + ok[1:], 11) is None          # virtual sample (TRANSPARENT edge)
    assert _schedule(dll, [[(0, 0.5), (0, 0.5)]] + ok[1:], 11) is None           # repeated sample (REPLICATE edge)
    assert _schedule(dll, ok[:5] + [[]] + ok[6:], 11) is None                    # empty window
    assert _schedule(dll, ok[:5] + [[(1, 1.0)]] + ok[6:], 11) is None            # window jumps back by more than a dropped tap
    wide = [[(s, 1.0 / 40) for s in range(o, o + 40)] for o in range(30)]
    assert _schedule(dll, wide, 70) is None                                      # more than 16 windows open at once
