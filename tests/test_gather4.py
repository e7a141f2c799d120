"""The gather-form streaming kernels (csrc/gather4*.cu*): images with a real alpha channel -- colour premultiplied by
alpha in the vertical pass (reference: src/imagew-main.c:880-885), un-premultiplied by the FP32-rounded resized alpha
and composited over a background in the horizontal epilogue (:1057-1061, :1091-1115), any output depth / dither /
colour count (:669-765).  Shapes cover RGBA and gray+alpha, 8- and 16-bit input, enlargements and mild reductions
(windows of 2, 4, 6 and 8 taps), alpha kept or replaced by a solid / checkerboard / translucent background, 8- and
16-bit output, none / ordered / random dither, posterisation, intermediate clamping, crops, ragged widths, bands.

All tests call through the C ABI and compare with the unmodified reference build, byte for byte.
"""
import ctypes as C

import numpy as np
import pytest

from imageworsener_b200 import configs, libimagew as L

RGBA, GA = L.IW_IMGTYPE_RGBA, L.IW_IMGTYPE_GRAYA

SHAPES = [
    # name, in_w, in_h, type, depth, out_w, out_h, arguments
    ("rgba16_up4_mitchell_bkgd16", 256, 256, RGBA, 16, 1024, 1024, "-filter mitchell -depth 16 -bkgd 369"),
    ("rgba8_up2_catrom_keepalpha", 300, 200, RGBA, 8, 600, 400, "-filter catrom"),
    ("rgba8_up3_lanczos_o", 200, 150, RGBA, 8, 600, 450, "-filter lanczos -dither o"),
    ("rgba8_down2_catrom_r", 800, 600, RGBA, 8, 400, 300, "-filter catrom -dither r"),
    ("rgba16_down15_mitchell_checker", 600, 450, RGBA, 16, 400, 300, "-filter mitchell -bkgd e42,0f5 -checkersize 7 -checkerorigin 2,3"),
    ("rgba8_translucent_bkgd", 320, 240, RGBA, 8, 400, 300, "-filter triangle -bkgd 00ff5550"),
    ("rgba8_translucent_checker", 320, 240, RGBA, 8, 400, 300, "-filter catrom -bkgd e42d,00ff5550 -checkersize 5"),
    ("ga8_up2_hermite", 301, 203, GA, 8, 602, 406, "-filter hermite -grayscale -dither halftone"),
    ("ga16_up_box_bkgd", 200, 160, GA, 16, 517, 333, "-filter box -grayscale -bkgd 808080 -depth 16"),
    ("ga8_down_triangle_cc", 500, 400, GA, 8, 333, 250, "-filter triangle -grayscale -cc 6 -dither o"),
    ("rgba8_ragged_r2", 333, 257, RGBA, 8, 417, 300, "-filter catrom -dither r2 -randseed 7"),
    ("rgba16_intclamp", 200, 200, RGBA, 16, 500, 500, "-filter lanczos -intclamp -depth 16"),
    ("rgba8_crop", 400, 300, RGBA, 8, 500, 380, "-filter mitchell -crop 13,7,301,211 -bkgd 123"),
    ("rgba8_depth5", 300, 300, RGBA, 8, 450, 450, "-filter catrom -depth 5,6,5,3 -dither o"),
    ("rgba8_samesize_bkgd", 300, 200, RGBA, 8, 300, 200, "-filter mitchell -bkgd 0a0"),
    ("rgba8_gaussian_w8", 300, 200, RGBA, 8, 420, 280, "-filter gaussian"),
    ("rgba8_negate", 200, 100, RGBA, 8, 300, 150, "-filter catrom -negate"),
    # 16-bit output without dither: the level model (csrc/estimate.h) settles most samples, the table the rest
    ("rgba16_model_rec709", 240, 200, RGBA, 16, 700, 555, "-filter catrom -depth 16 -bkgd 2a7 -cs rec709"),
    ("rgba8_model_to16", 300, 220, RGBA, 8, 640, 480, "-filter lanczos -depth 16 -bkgd fff"),
    ("rgba16_model_depth12", 256, 256, RGBA, 16, 600, 600, "-filter mitchell -depth 12 -bkgd 000"),
    ("ga16_model_down", 640, 480, GA, 16, 500, 377, "-filter mitchell -grayscale -depth 16 -bkgd 777"),
    ("rgba16_model_gamma_falls_back", 200, 200, RGBA, 16, 420, 420, "-filter mitchell -depth 16 -bkgd 369 -cs gamma1.8"),
]


# windows of more than 8 taps: the tile kernels (csrc/gather5.cu)
TILE_SHAPES = [
    ("t_rgba8_thumb_auto", 480, 270, RGBA, 8, 64, 36, ""),
    ("t_rgba8_down2_lanczos", 400, 300, RGBA, 8, 200, 150, "-filter lanczos"),
    ("t_rgba8_down3_catrom_bkgd", 600, 420, RGBA, 8, 200, 140, "-filter catrom -bkgd 3a7"),
    ("t_rgba16_down25_mitchell_16", 500, 375, RGBA, 16, 200, 150, "-filter mitchell -depth 16 -bkgd 123"),
    ("t_rgba16_down4_lanczos_o", 640, 480, RGBA, 16, 160, 120, "-filter lanczos -dither o"),
    ("t_ga8_down5_keepalpha", 500, 400, GA, 8, 100, 80, "-filter catrom -grayscale"),
    ("t_ga16_down3_halftone", 420, 300, GA, 16, 140, 100, "-filter lanczos2 -grayscale -bkgd 888 -dither halftone"),
    ("t_rgba8_checker", 512, 384, RGBA, 8, 150, 113, "-filter mitchell -bkgd e42,0f5 -checkersize 9"),
    ("t_rgba8_translucent", 512, 384, RGBA, 8, 170, 128, "-filter catrom -bkgd 00ff5550"),
    ("t_rgba8_aniso", 512, 200, RGBA, 8, 100, 190, "-filter lanczos"),
    ("t_rgba8_crop_cc", 512, 384, RGBA, 8, 120, 90, "-filter catrom -crop 17,9,400,300 -cc 6 -dither o"),
    ("t_rgba8_intclamp", 400, 300, RGBA, 8, 130, 100, "-filter lanczos -intclamp"),
    ("t_rgba8_up_lanczos3x", 60, 40, RGBA, 8, 300, 200, "-filter lanczos"),
    ("t_rgba8_down12_box", 1200, 900, RGBA, 8, 100, 75, "-filter box"),
    ("t_rgba8_ragged", 333, 257, RGBA, 8, 77, 61, "-filter catrom -bkgd fff"),
    # images without alpha that the scatter-form kernels do not take: 16-bit output, more than 16 open windows
    ("t_rgb16_down2_to16", 400, 300, L.IW_IMGTYPE_RGB, 16, 200, 150, "-filter lanczos -depth 16"),
    ("t_rgb8_up_to16", 200, 150, L.IW_IMGTYPE_RGB, 8, 500, 377, "-filter mitchell -depth 16"),
    ("t_gray16_down3_to16", 600, 420, L.IW_IMGTYPE_GRAY, 16, 200, 140, "-filter catrom -grayscale -depth 16"),
    ("t_gray16_to16_rec709_o", 320, 240, L.IW_IMGTYPE_GRAY, 16, 251, 188, "-filter lanczos2 -grayscale -depth 16 -cs rec709 -dither o"),
    ("t_rgb8_up8_catrom", 60, 45, L.IW_IMGTYPE_RGB, 8, 480, 360, "-filter catrom -dither o"),
    ("t_rgb16_depth12", 300, 200, L.IW_IMGTYPE_RGB, 16, 260, 170, "-filter mitchell -depth 12"),
]


def _cfg(shape):
    name, iw, ih, typ, depth, ow, oh, extra = shape
    return configs.Config(name, iw, ih, typ, depth, ow, oh, f"-width {ow} -height {oh} {extra}")


def _rows_with_transparency(prod, cfg, seed=3):
    """Noise, with patches of fully transparent and fully opaque pixels (exact alpha 0 and max exercise the
    `alpha != 0` test and the transparent -> black rule)."""
    rows = configs.synth_rows(prod, cfg).copy()
    nch = L.imgtype_num_channels(cfg.imgtype)
    bps = cfg.bit_depth // 8
    px = rows[:, : cfg.in_w * nch * bps].reshape(cfg.in_h, cfg.in_w, nch, bps)
    px[: cfg.in_h // 3, : cfg.in_w // 2, nch - 1, :] = 0
    px[cfg.in_h // 2:, cfg.in_w // 3:, nch - 1, :] = 255
    # black and saturated colour (exact zeros / ones after the division by alpha), with and without opacity
    px[cfg.in_h // 4: cfg.in_h // 2, : cfg.in_w // 4, : nch - 1, :] = 0
    px[cfg.in_h // 2: 3 * cfg.in_h // 4, cfg.in_w // 2:, : nch - 1, :] = 255
    return rows


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + TILE_SHAPES, ids=[s[0] for s in SHAPES + TILE_SHAPES])
def test_gather_path_matches_reference(prod, ref, shape):
    cfg = _cfg(shape)
    rows = _rows_with_transparency(prod, cfg)
    ia, a = configs.run_host(prod, cfg, rows)
    ib, b = configs.run_host(ref, cfg, rows)
    assert (ia.width, ia.height, ia.bit_depth, ia.imgtype, ia.bpr) == (ib.width, ib.height, ib.bit_depth, ib.imgtype, ib.bpr)
    nd = int(np.count_nonzero(a != b))
    assert nd == 0, f"{nd} differing bytes of {a.size}"
    # and the generic kernels agree (the option switches the gather path off)
    _, c = configs.run_host(prod, cfg, rows, options={"cuda:gather": 0})
    assert np.array_equal(c, b)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES[:12], ids=[s[0] for s in SHAPES[:12]])
def test_these_jobs_are_planned_onto_the_gather_kernels(prod, shape):
    from test_intermediate import Info, _job_api
    cfg = _cfg(shape)
    rows = configs.synth_rows(prod, cfg)
    dll = prod.dll
    _job_api(dll)
    ctx = prod.iw_create_context(None)
    prod.set_input_from_numpy(ctx, rows, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    cfg.apply(prod, ctx)
    job = dll.iwb200_job_create(ctx, 0)
    assert job, prod.errmsg(ctx)
    info = Info()
    dll.iwb200_job_get_info(job, C.byref(info))
    dll.iwb200_job_destroy(job)
    prod.iw_destroy_context(ctx)
    assert info.uses_fused_kernel == 4


@pytest.mark.gpu
@pytest.mark.parametrize("shape", TILE_SHAPES[:12] + TILE_SHAPES[15:], ids=[s[0] for s in TILE_SHAPES[:12] + TILE_SHAPES[15:]])
def test_these_jobs_are_planned_onto_the_tile_kernels(prod, shape):
    from test_intermediate import Info, _job_api
    cfg = _cfg(shape)
    rows = configs.synth_rows(prod, cfg)
    dll = prod.dll
    _job_api(dll)
    ctx = prod.iw_create_context(None)
    prod.set_input_from_numpy(ctx, rows, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    cfg.apply(prod, ctx)
    job = dll.iwb200_job_create(ctx, 0)
    assert job, prod.errmsg(ctx)
    info = Info()
    dll.iwb200_job_get_info(job, C.byref(info))
    dll.iwb200_job_destroy(job)
    prod.iw_destroy_context(ctx)
    assert info.uses_fused_kernel == 5


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[3], SHAPES[7], TILE_SHAPES[0], TILE_SHAPES[3], TILE_SHAPES[5]], ids=lambda s: s[0])
@pytest.mark.parametrize("nbands", [2, 5])
def test_gather_path_bands_concatenate_to_reference(prod, ref, shape, nbands):
    cfg = _cfg(shape)
    rows = _rows_with_transparency(prod, cfg)
    parts = [configs.run_host(prod, cfg, rows, options={"cuda:band": f"{i}/{nbands}"})[1] for i in range(nbands)]
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(np.concatenate(parts, axis=0), b)


@pytest.mark.gpu
def test_cfg3_quarter_scale(prod, ref):
    """BASELINE configuration 3 (RGBA16 4x enlargement, Mitchell, premultiply + solid background, 16-bit out) at 1/4 size."""
    cfg = configs.get("cfg3", scale=4)
    rows = configs.synth_rows(prod, cfg)
    _, a = configs.run_host(prod, cfg, rows)
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [8, 16])
def test_fresh_jobs_in_bands_repeatedly(prod, ref, depth):
    """Regression: a band's pipeline ends with stages that lie outside the band's slab of input rows; the copy engine
    zero-fills those within a few cycles, which once overtook the last shared-memory reads of the slot being refilled
    (six wrong output rows in one warp column, a few times in ten runs, only with freshly built jobs).  Every job here
    is built from scratch (cuda:plan_cache=0)."""
    cfg = configs.Config("bands", 300, 260, RGBA, depth, 450, 390, "-width 450 -height 390 -filter mitchell -depth 16 -bkgd e42")
    rows = configs.synth_rows(prod, cfg)
    _, b = configs.run_host(ref, cfg, rows)
    for it in range(12):
        parts = [configs.run_host(prod, cfg, rows, options={"cuda:band": f"{i}/5", "cuda:plan_cache": 0})[1] for i in range(5)]
        assert np.array_equal(np.concatenate(parts, axis=0), b), f"iteration {it}"
        _, a = configs.run_host(prod, cfg, rows, options={"cuda:devices": "0,0,0,0,0", "cuda:plan_cache": 0})
        assert np.array_equal(a, b), f"iteration {it} (threads)"
