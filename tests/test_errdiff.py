"""Error-diffusion dither on the device (csrc/errdiff.cuh: one warp per channel, the chain in one lane, error rows in
shared memory) against the reference (src/imagew-main.c:497-567, :690-765, :1124-1136): all eight diffusion kernels,
8- and 16-bit output, gray / RGB / RGBA with background, posterised levels, linear output space (the quantiser's
non-table mode), widths of 1-3 pixels (every neighbour out of range) and chunked processing (error rows carried across
kernel launches)."""
import numpy as np
import pytest

from imageworsener_b200 import configs, libimagew as L

RGB, GRAY, RGBA = L.IW_IMGTYPE_RGB, L.IW_IMGTYPE_GRAY, L.IW_IMGTYPE_RGBA
KERNELS = ["f", "jjn", "stucki", "burkes", "sierra", "sierra2", "sierralite", "atkinson"]

CASES = [(f"rgb8_{k}", 400, 300, RGB, 8, 233, 171, f"-filter catrom -dither {k}") for k in KERNELS] + [
    ("gray16_f", 300, 200, GRAY, 16, 450, 260, "-filter mitchell -grayscale -dither f -depth 16"),
    ("gray8_depth5_jjn", 300, 200, GRAY, 8, 200, 133, "-grayscale -dither jjn -depth 5"),
    ("rgb8_cc_stucki", 300, 200, RGB, 8, 260, 170, "-dither stucki -cc 4"),
    ("rgb8_linear_f", 300, 200, RGB, 8, 260, 170, "-dither f -cs linear"),
    ("rgba8_bkgd_atkinson", 200, 160, RGBA, 8, 300, 240, "-filter catrom -dither atkinson -bkgd 4a8"),
    ("rgba8_keepalpha_f", 200, 160, RGBA, 8, 150, 120, "-dither f"),
    ("w1", 30, 200, RGB, 8, 1, 150, "-dither f"),
    ("w2", 30, 200, RGB, 8, 2, 150, "-dither jjn"),
    ("w3", 30, 200, GRAY, 8, 3, 150, "-grayscale -dither sierra"),
    ("mixed_dithers", 300, 200, RGB, 8, 260, 170, "-ditherred f -dithergreen o -ditherblue jjn"),
    ("wide", 6000, 40, RGB, 8, 5000, 30, "-filter box -dither f"),
]


def _cfg(case):
    name, iw, ih, typ, depth, ow, oh, extra = case
    return configs.Config(name, iw, ih, typ, depth, ow, oh, f"-width {ow} -height {oh} {extra}")


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_error_diffusion_matches_reference(prod, ref, case):
    cfg = _cfg(case)
    rows = configs.synth_rows(prod, cfg)
    ia, a = configs.run_host(prod, cfg, rows)
    ib, b = configs.run_host(ref, cfg, rows)
    assert (ia.width, ia.height, ia.bit_depth, ia.imgtype, ia.bpr) == (ib.width, ib.height, ib.bit_depth, ib.imgtype, ib.bpr)
    nd = int(np.count_nonzero(a != b))
    assert nd == 0, f"{nd} differing bytes of {a.size}"
    # the per-thread kernel agrees (cuda:errdiff_generic selects it)
    _, c = configs.run_host(prod, cfg, rows, options={"cuda:errdiff_generic": 1})
    assert np.array_equal(c, b)


@pytest.mark.gpu
def test_error_diffusion_smooth_image(prod, ref):
    """A gradient: long runs of samples near level boundaries and exactly on them (the `exact` path spreads no error)."""
    w, h = 640, 200
    yy, xx = np.mgrid[0:h, 0:w * 3]
    rows = ((xx // 3 * 255) // (w - 1)).astype(np.uint8)
    rows[h // 2:] = ((yy[h // 2:] * 3) % 256).astype(np.uint8)
    cfg = configs.Config("grad", w, h, RGB, 8, 320, 100, "-width 320 -height 100 -filter triangle -dither f -depth 3")
    _, a = configs.run_host(prod, cfg, rows)
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(a, b)
