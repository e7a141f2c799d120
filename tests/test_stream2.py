"""The streaming fast path (csrc/stream2*.cu, csrc/fused.cuh): every code path must give the reference's bytes.

The fast path has three implementations that a job can land on -- generation-2 kernels with the uniform
(straight-line) path, generation-2 kernels forced onto the general path, and the generation-1 kernels (used when
the caller's rows are not 16-byte aligned, or on request) -- and the generic per-sample kernels behind them.  The
shapes below are chosen to hit: integer and non-integer scale factors (uniform / general), enlargement, gray and
RGB, no / ordered / halftone / random dither, widths whose row stride is not a multiple of 16 (generation-1
fallback), images too small for a uniform range, and output-row bands (strips that start in the middle of the
image, absolute dither and PRNG coordinates).

All tests call through the C ABI (iw_process_image) and compare with the unmodified reference build.
"""
import numpy as np
import pytest

from imageworsener_b200 import configs, libimagew as L

RGB, GRAY = L.IW_IMGTYPE_RGB, L.IW_IMGTYPE_GRAY

SHAPES = [
    # name, in_w, in_h, type, out_w, out_h, arguments
    ("down2_lanczos_o", 1024, 768, RGB, 512, 384, "-filter lanczos -dither o"),
    ("down2_catrom_r", 1024, 768, RGB, 512, 384, "-filter catrom -dither r"),
    ("down2_mitchell_gray", 1024, 768, GRAY, 512, 384, "-filter mitchell"),
    ("down3_lanczos2_halftone", 1536, 960, RGB, 512, 320, "-filter lanczos2 -dither halftone"),
    ("down4_cubic", 2048, 512, RGB, 512, 128, "-filter cubic0.5,0.5 -dither o"),
    ("down2x_only", 1280, 400, RGB, 640, 400, "-filter lanczos -dither o"),
    ("odd_ratio", 1000, 700, RGB, 617, 333, "-filter lanczos -dither o"),
    ("odd_ratio_gray_r", 1001, 701, GRAY, 333, 233, "-filter catrom -dither r"),
    ("up_hermite", 320, 240, RGB, 512, 400, "-filter hermite -dither halftone"),
    ("up_gray", 256, 256, GRAY, 640, 512, "-filter mitchell -dither o"),
    ("wide_short", 4112, 64, RGB, 2056, 32, "-filter lanczos2"),
    ("tiny", 48, 40, RGB, 24, 20, "-filter lanczos -dither o"),
    ("unaligned_stride", 1022, 300, RGB, 511, 150, "-filter lanczos -dither o"),
    ("box_7p5", 1920, 1080, RGB, 256, 144, "-filter box"),
    ("down2_lanczos4", 1024, 512, RGB, 512, 256, "-filter lanczos4 -dither o"),        # 8 open windows, step 2
    ("down4_hermite", 2048, 512, GRAY, 512, 128, "-filter hermite -dither r"),           # 2 open windows, step 4
    ("down4_quadratic", 2048, 512, RGB, 512, 128, "-filter quadratic"),                  # 3 open windows, step 4
    ("down2_triangle", 1024, 512, RGB, 512, 256, "-filter triangle -dither halftone"),   # 2 open windows, step 2
    # windows with holes: taps whose weight is exactly zero are absent from the reference's lists
    ("down3_catrom_holes", 1536, 960, RGB, 512, 320, "-filter catrom"),
    ("down3_catrom_holes_r", 1536, 480, GRAY, 512, 160, "-filter catrom -grayscale -dither r"),
    ("down5_mitchell_holes", 2560, 640, RGB, 512, 128, "-filter mitchell -dither o"),
    ("up2_catrom_holes", 256, 200, RGB, 512, 400, "-filter catrom -dither o"),
    # enlargements: 12 and 16 open windows (general-only kernels)
    ("up2_lanczos_s12", 320, 200, RGB, 640, 400, "-filter lanczos -dither o"),
    ("up3_catrom_s12", 200, 150, RGB, 600, 450, "-filter catrom -dither r"),
    ("up4_mitchell_s16", 160, 120, RGB, 640, 480, "-filter mitchell"),
    ("up2p4_lanczos_s16_gray", 200, 160, GRAY, 480, 384, "-filter lanczos -grayscale -dither halftone"),
    ("up4_down2_mixed", 160, 800, RGB, 640, 400, "-filter mitchell -dither o"),          # 16 windows across, 4 down
    # gray input written as RGB (no -grayscale): one intermediate plane, three output bytes
    ("gray_as_rgb_o", 1024, 512, GRAY, 512, 256, "-filter lanczos -dither o"),
    ("gray_as_rgb_r", 1024, 512, GRAY, 512, 256, "-filter catrom -dither r"),                # three different draw sequences
    ("gray_as_rgb_r2", 1000, 500, GRAY, 333, 250, "-filter catrom -dither r2"),              # same pattern on the three
    ("gray_as_rgb_none_odd", 1000, 500, GRAY, 617, 333, "-filter lanczos2"),
    ("gray_as_rgb_up_s12", 200, 160, GRAY, 400, 320, "-filter lanczos -dither halftone"),
]


def _cfg(shape):
    name, iw, ih, typ, ow, oh, extra = shape
    return configs.Config(name, iw, ih, typ, 8, ow, oh, f"-width {ow} -height {oh} {extra}")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
@pytest.mark.parametrize("variant", ["gen2", "gen2-general", "gen1"])
def test_fast_path_variants_match_reference(prod, ref, shape, variant):
    cfg = _cfg(shape)
    rows = configs.synth_rows(prod, cfg)
    opts = {"gen2": {}, "gen2-general": {"cuda:uniform": 0}, "gen1": {"cuda:gen": 1}}[variant]
    ia, a = configs.run_host(prod, cfg, rows, options=opts)
    ib, b = configs.run_host(ref, cfg, rows)
    assert (ia.width, ia.height, ia.bit_depth, ia.imgtype, ia.bpr) == (ib.width, ib.height, ib.bit_depth, ib.imgtype, ib.bpr)
    nd = int(np.count_nonzero(a != b))
    assert nd == 0, f"{nd} differing bytes of {a.size}"


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[1], SHAPES[6]], ids=lambda s: s[0])
@pytest.mark.parametrize("nbands", [2, 5])
def test_fast_path_bands_concatenate_to_reference(prod, ref, shape, nbands):
    cfg = _cfg(shape)
    rows = configs.synth_rows(prod, cfg)
    parts = [configs.run_host(prod, cfg, rows, options={"cuda:band": f"{i}/{nbands}"})[1] for i in range(nbands)]
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(np.concatenate(parts, axis=0), b)


@pytest.mark.gpu
@pytest.mark.parametrize("strip,seg", [(24, 20), (56, 36), (200, 100)])
def test_fast_path_partition_sizes_do_not_matter(prod, ref, strip, seg):
    """Strip / segment lengths only change which CTA computes what: any choice must give the same bytes."""
    cfg = _cfg(SHAPES[0])
    rows = configs.synth_rows(prod, cfg)
    _, a = configs.run_host(prod, cfg, rows, options={"cuda:strip2": strip, "cuda:seg2": seg})
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("w,typ", [(1001, GRAY), (333, RGB), (1022, RGB)])
def test_fast_path_padded_rows_with_ragged_last_columns(prod, ref, w, typ):
    """Rows padded to a multiple of 16 bytes (so the TMA kernels are used) while the number of byte columns is not a
    multiple of 4 or 128: the last warp's tile is partly outside the tensor and its last thread stores fewer than
    four columns."""
    nch = L.imgtype_num_channels(typ)
    h, ow, oh = 300, w // 2, 150
    cfg = configs.Config("padded", w, h, typ, 8, ow, oh, f"-width {ow} -height {oh} -filter lanczos -dither o")
    tight = configs.synth_rows(prod, cfg)
    bpr = (w * nch + 15) // 16 * 16 + 16
    rows = np.zeros((h, bpr), np.uint8)
    rows[:, : w * nch] = tight
    rows[:, w * nch:] = 0xA5                      # padding bytes must never influence the result
    ia, a = configs.run_host(prod, cfg, rows)
    ib, b = configs.run_host(ref, cfg, tight)
    assert np.array_equal(a, b)


def _near_integer_case(seed):
    """8-bit jobs the TMA kernels accept, with scale factors at or next to small integers (where uniform ranges
    break or windows outgrow a cycle), random filters / dithers / sizes, rows padded to 16 bytes."""
    rng = np.random.default_rng(seed)
    typ = RGB if rng.random() < 0.6 else GRAY
    flt = rng.choice(["lanczos", "lanczos2", "lanczos4", "catrom", "mitchell", "hermite", "triangle", "quadratic", "gaussian",
                      "bspline", "hann3", "blackman", "box", "mix"])
    dither = rng.choice(["", "-dither o", "-dither halftone", "-dither r", "-dither r2"])
    fx, fy = rng.choice([1.5, 2, 3, 4], 2)
    ow, oh = int(rng.integers(90, 400)), int(rng.integers(40, 160))
    iw = max(8, int(round(ow * fx)) + int(rng.integers(-3, 4)))
    ih = max(8, int(round(oh * fy)) + int(rng.integers(-3, 4)))
    blur = f" -blur {rng.choice([1.0, 1.2, 0.9])}" if rng.random() < 0.25 else ""
    return configs.Config(f"near{seed}", iw, ih, typ, 8, ow, oh, f"-width {ow} -height {oh} -filter {flt} {dither}{blur}")


def _enlargement_case(seed):
    """Enlargements at or next to 1.5x .. 4x per axis (independently; one axis may shrink): 12 and 16 open windows,
    end taps dropped as exact zeros, gray input written as RGB, rows padded to 16 bytes."""
    rng = np.random.default_rng(seed)
    typ = RGB if rng.random() < 0.6 else GRAY
    flt = rng.choice(["lanczos", "lanczos2", "catrom", "mitchell", "hermite", "triangle", "quadratic", "gaussian", "bspline", "mix"])
    dither = rng.choice(["", "-dither o", "-dither halftone", "-dither r", "-dither r2"])
    fx, fy = rng.choice([1.5, 2, 2.5, 3, 4, 0.5], 2)
    iw, ih = int(rng.integers(40, 160)), int(rng.integers(24, 90))
    ow = max(8, int(round(iw * fx)) + int(rng.integers(-3, 4)))
    oh = max(8, int(round(ih * fy)) + int(rng.integers(-3, 4)))
    gray = " -grayscale" if typ == GRAY and rng.random() < 0.5 else ""
    return configs.Config(f"up{seed}", iw, ih, typ, 8, ow, oh, f"-width {ow} -height {oh} -filter {flt}{gray} {dither}")


@pytest.mark.gpu
@pytest.mark.parametrize("gen", [_near_integer_case, _enlargement_case], ids=["reduce", "enlarge"])
@pytest.mark.parametrize("seed", range(80))
def test_fast_path_near_integer_factors_match_reference(prod, ref, seed, gen):
    cfg = gen(seed)
    nch = L.imgtype_num_channels(cfg.imgtype)
    tight = configs.synth_rows(prod, cfg, seed=0x1234 + seed)
    bpr = (cfg.in_w * nch + 15) // 16 * 16
    rows = np.full((cfg.in_h, bpr), 0x5A, np.uint8)
    rows[:, : cfg.in_w * nch] = tight
    _, b = configs.run_host(ref, cfg, tight)
    for opts in ({}, {"cuda:uniform": 0}):
        _, a = configs.run_host(prod, cfg, rows, options=opts)
        nd = int(np.count_nonzero(a != b))
        assert nd == 0, f"{cfg.args!r} {cfg.in_w}x{cfg.in_h} {opts}: {nd} differing bytes of {a.size}"


@pytest.mark.gpu
@pytest.mark.parametrize("extra,opts", [
    ("-filter lanczos -dither o", {"cuda:chunk_mb": 1}),          # many chunks of output rows per pass
    ("-filter catrom -dither r", {"cuda:chunk_mb": 1}),           # ... with absolute PRNG coordinates
    ("-filter lanczos -dither o -negate", {}),                    # post kernel behind the horizontal pass
    ("-filter lanczos -dither o -crop 32,10,800,500", {}),        # crop origin on a 16-byte boundary (TMA kernels)
    ("-filter lanczos -dither o -crop 33,11,801,499", {}),        # ... and off it: the host copy re-bases the crop on an aligned row
    ("-filter lanczos -dither o -crop 33,11,801,499", {"cuda:keep_pitch": 1}),   # ... or not (generation-1 kernels)
    ("-filter mitchell -crop 64,0,640,768", {"cuda:chunk_mb": 1}),
    ("-filter catrom -dither r -crop 5,3,1000,700", {"cuda:band": "1/3"}),       # a band of a cropped image
], ids=["chunks", "chunks-random", "negate", "crop-aligned", "crop-unaligned", "crop-unaligned-gen1", "crop-chunks", "crop-band"])
@pytest.mark.parametrize("depth", [8, 16])
def test_fast_path_chunks_negate_crop(prod, ref, extra, opts, depth):
    ow, oh = (512, 384) if "-crop" not in extra else (400, 250)
    cfg = configs.Config("misc", 1024, 768, RGB, depth, ow, oh, f"-width {ow} -height {oh} {extra}")
    rows = configs.synth_rows(prod, cfg)
    if "cuda:band" in opts:                                      # all three bands, concatenated
        a = np.concatenate([configs.run_host(prod, cfg, rows, options={"cuda:band": f"{i}/3"})[1] for i in range(3)], axis=0)
    else:
        _, a = configs.run_host(prod, cfg, rows, options=opts)
    _, b = configs.run_host(ref, cfg, rows)
    nd = int(np.count_nonzero(a != b))
    assert nd == 0, f"{nd} differing bytes of {a.size}"


SHAPES16 = [
    ("gray16_down2_catrom_r", 1024, 768, GRAY, 512, 384, "-filter catrom -grayscale -dither r"),
    ("gray16_down2_lanczos_o", 1024, 768, GRAY, 512, 384, "-filter lanczos -grayscale -dither o"),
    ("rgb16_down2_lanczos_o", 1024, 512, RGB, 512, 256, "-filter lanczos -dither o"),
    ("rgb16_odd_mitchell", 1000, 512, RGB, 617, 300, "-filter mitchell"),
    ("gray16_up_hermite", 320, 240, GRAY, 512, 400, "-filter hermite -grayscale -dither halftone"),
    ("gray16_unaligned_stride", 1001, 400, GRAY, 500, 200, "-filter catrom -grayscale -dither r"),     # generic kernels
    ("rgb16_down3_lanczos2", 1536, 480, RGB, 512, 160, "-filter lanczos2 -dither r"),
    ("rgb16_up2_lanczos_s12", 320, 200, RGB, 640, 400, "-filter lanczos -dither o"),
    ("gray16_up4_mitchell_s16", 160, 120, GRAY, 640, 480, "-filter mitchell -grayscale -dither r"),
    ("gray16_as_rgb_r", 1024, 512, GRAY, 512, 256, "-filter catrom -dither r"),
    ("gray16_as_rgb_o", 640, 400, GRAY, 320, 200, "-filter lanczos -dither o"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES16, ids=[s[0] for s in SHAPES16])
@pytest.mark.parametrize("variant", ["gen2", "gen2-general", "generic"])
def test_fast_path_16bit_input_matches_reference(prod, ref, shape, variant):
    """16-bit big-endian input (8-bit output): the vertical kernel's 16-bit variant (global 65 536-entry table) with
    the same horizontal kernels; an input gamma as in BASELINE cfg5 for the gray cases."""
    name, iw, ih, typ, ow, oh, extra = shape
    cfg = configs.Config(name, iw, ih, typ, 16, ow, oh, f"-width {ow} -height {oh} {extra}",
                         input_cs=("gamma", 2.2) if typ == GRAY else None)
    rows = configs.synth_rows(prod, cfg)
    opts = {"gen2": {}, "gen2-general": {"cuda:uniform": 0}, "generic": {"cuda:fused": 0}}[variant]
    ia, a = configs.run_host(prod, cfg, rows, options=opts)
    ib, b = configs.run_host(ref, cfg, rows)
    assert (ia.width, ia.height, ia.bit_depth, ia.imgtype, ia.bpr) == (ib.width, ib.height, ib.bit_depth, ib.imgtype, ib.bpr)
    nd = int(np.count_nonzero(a != b))
    assert nd == 0, f"{nd} differing bytes of {a.size}"


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [{}, {"cuda:keep_pitch": 1}], ids=["aligned-device-copy", "caller-pitch"])
@pytest.mark.parametrize("depth", [8, 16])
def test_fast_path_unaligned_host_rows(prod, ref, opts, depth):
    """Host rows whose stride is not a multiple of 16 bytes: iw_process_image gives the device copy an aligned pitch
    (2-D copy), so the TMA kernels run; with cuda:keep_pitch the device copy keeps the caller's stride and the
    job falls back (generation 1 for 8-bit, generic kernels for 16-bit).  Same bytes either way."""
    cfg = configs.Config("unal", 1001, 400, RGB, depth, 500, 200, "-width 500 -height 200 -filter lanczos -dither o")
    rows = configs.synth_rows(prod, cfg)
    assert rows.shape[1] % 16 != 0
    _, a = configs.run_host(prod, cfg, rows, options=opts)
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [s for s in SHAPES if "holes" in s[0] or "_s1" in s[0] or "gray_as_rgb" in s[0]], ids=lambda s: s[0])
def test_windows_with_holes_take_the_streaming_kernels(prod, shape):
    """Jobs whose tap lists skip zero-weight samples, and enlargements with up to 16 open windows, are planned onto
    the generation-2 kernels, not the older ones."""
    import ctypes as C
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
    assert info.uses_fused_kernel == 2
