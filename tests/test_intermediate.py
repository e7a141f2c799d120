"""Intermediate-sample parity (north star: "intermediate linear samples within 1e-12 relative").

The reference keeps the vertical-pass result as FP32 (src/imagew-main.c:905-912) and does not expose it; the
oracle restatement (oracle/iw_oracle.c, pinned against the reference build and its golden vectors) keeps the plane
of the last processed channel.  The CUDA library exposes its FP32 intermediate through the verification hook
iwb200_job_read_intermediate().  Because the arithmetic order is fixed and the result is rounded to float, the two
must agree BIT FOR BIT, which is what is asserted (0 mismatches; the 1e-12 bar is implied).
"""
import ctypes as C
import os

import numpy as np
import pytest

from imageworsener_b200 import configs, libimagew as L

RGB, GRAY = L.IW_IMGTYPE_RGB, L.IW_IMGTYPE_GRAY
CASES = [
    ("down2_lanczos_rgb", 512, 384, RGB, 8, 256, 192, "-filter lanczos -dither o", {}),
    ("down2_lanczos_rgb_general", 512, 384, RGB, 8, 256, 192, "-filter lanczos -dither o", {"cuda:uniform": 0}),
    ("down2_lanczos_rgb_gen1", 512, 384, RGB, 8, 256, 192, "-filter lanczos -dither o", {"cuda:gen": 1}),
    ("down2_lanczos_rgb_gen3", 512, 384, RGB, 8, 256, 192, "-filter lanczos -dither o", {"cuda:gen": 3}),
    ("rgba16_up_gather", 200, 150, L.IW_IMGTYPE_RGBA, 16, 500, 380, "-filter mitchell -bkgd 369", {}),
    ("ga8_down_gather", 400, 300, L.IW_IMGTYPE_GRAYA, 8, 260, 200, "-filter catrom", {}),
    ("rgba8_down3_tile", 420, 300, L.IW_IMGTYPE_RGBA, 8, 140, 100, "-filter catrom -bkgd 246", {}),
    ("ga16_down4_tile", 400, 320, L.IW_IMGTYPE_GRAYA, 16, 100, 80, "-filter lanczos", {}),
    ("down2_catrom_gray_gen3", 1002, 702, GRAY, 8, 501, 351, "-filter catrom -grayscale -dither r", {"cuda:gen": 3}),
    ("odd_catrom_gray", 500, 333, GRAY, 8, 217, 140, "-filter catrom -dither r", {}),
    ("up_mitchell_rgb", 200, 150, RGB, 8, 320, 260, "-filter mitchell", {}),
    ("gray16_generic", 300, 200, GRAY, 16, 150, 100, "-filter catrom -dither r", {}),
    ("gray16_stream", 320, 200, GRAY, 16, 160, 100, "-filter catrom -grayscale -dither r", {}),
    ("rgb16_stream", 320, 200, RGB, 16, 150, 101, "-filter lanczos -dither o", {}),
    # cubic filters at odd integer factors: the reference drops the exactly-zero taps at |x| = 1, the streaming
    # schedules carry a zero weight there instead (pipeline.cu:build_sched)
    ("down3_catrom_holes", 480, 360, RGB, 8, 160, 120, "-filter catrom -dither o", {}),
    ("down3_catrom_holes_general", 480, 360, RGB, 8, 160, 120, "-filter catrom -dither o", {"cuda:uniform": 0}),
    ("down3_catrom_holes_gen1", 480, 360, RGB, 8, 160, 120, "-filter catrom -dither o", {"cuda:gen": 1}),
    ("down5_mitchell_holes_gray16", 640, 400, GRAY, 16, 128, 80, "-filter mitchell -grayscale", {}),
    ("up2_lanczos_s12", 200, 120, RGB, 8, 400, 240, "-filter lanczos -dither o", {}),
    ("up4_mitchell_s16_gray16", 100, 80, GRAY, 16, 400, 320, "-filter mitchell -grayscale -dither r", {}),
    ("gray_as_rgb", 512, 384, GRAY, 8, 256, 192, "-filter lanczos -dither r", {}),
]


class Info(C.Structure):
    _fields_ = [("out_width", C.c_int), ("out_height", C.c_int), ("out_imgtype", C.c_int),
                ("out_bit_depth", C.c_int), ("out_sampletype", C.c_int), ("out_channels", C.c_int),
                ("out_bpr", C.c_size_t), ("in_width_phys", C.c_int), ("in_height_phys", C.c_int),
                ("in_bpr", C.c_size_t), ("bandable", C.c_int), ("uses_fused_kernel", C.c_int),
                ("taps_v", C.c_double), ("taps_h", C.c_double)]


def _job_api(dll):
    dll.iwb200_job_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
    dll.iwb200_job_create.restype = C.c_void_p
    dll.iwb200_job_create.argtypes = [C.c_void_p, C.c_int]
    dll.iwb200_job_destroy.argtypes = [C.c_void_p]
    dll.iwb200_job_run_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    dll.iwb200_job_last_error.restype = C.c_char_p
    dll.iwb200_job_last_error.argtypes = [C.c_void_p]
    dll.iwb200_job_read_intermediate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int),
                                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]


def _near_cases():
    from test_stream2 import _near_integer_case
    out = []
    for seed in range(100, 124):
        c = _near_integer_case(seed)
        out.append((c.name, c.in_w, c.in_h, c.imgtype, 8, c.out_w, c.out_h, c.args.split(f"-height {c.out_h} ", 1)[1], {}, True))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c + (False,) for c in CASES] + _near_cases(), ids=lambda c: c[0])
def test_fp32_intermediate_is_bit_identical_to_oracle(prod, oracle_port, case):
    name, iw, ih, typ, depth, ow, oh, extra, opts, pad = case
    cfg = configs.Config(name, iw, ih, typ, depth, ow, oh, f"-width {ow} -height {oh} {extra}")
    rows = configs.synth_rows(prod, cfg)
    if pad:                                      # rows padded to 16 bytes so that the TMA kernels take the job
        nchp = L.imgtype_num_channels(typ)
        padded = np.full((ih, (iw * nchp + 15) // 16 * 16), 0x33, np.uint8)
        padded[:, : iw * nchp] = rows
        tight, rows = rows, padded
    else:
        tight = rows
    _check_planes(prod, cfg, rows, tight, opts)


def _check_planes(prod, cfg, rows, tight, opts):
    iw, ih, ow, oh = cfg.in_w, cfg.in_h, cfg.out_w, cfg.out_h
    # oracle: last processed channel's plane
    from oracle_port import OraclePort
    olib = OraclePort()
    octx = olib.iw_create_context(None)
    olib.set_input_from_numpy(octx, tight, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    cfg.apply(olib, octx)
    assert olib.iw_process_image(octx), olib.errmsg(octx)
    want = olib.intermediate(octx)
    _, opx = olib.get_output_numpy(octx)

    # product, through the job layer so the intermediate can be read back before the job is released
    dll = prod.dll
    _job_api(dll)
    ctx = prod.iw_create_context(None)
    prod.set_input_from_numpy(ctx, rows, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    for k, v in opts.items():
        prod.iw_set_option(ctx, k.encode(), str(v).encode())
    cfg.apply(prod, ctx)
    job = dll.iwb200_job_create(ctx, 0)
    assert job, prod.errmsg(ctx)
    nch = L.imgtype_num_channels(cfg.imgtype)
    info = Info()
    dll.iwb200_job_get_info(job, C.byref(info))
    assert (info.out_width, info.out_height) == (ow, oh)
    out = np.zeros((oh, info.out_bpr), np.uint8)
    src = np.ascontiguousarray(rows)
    ok = dll.iwb200_job_run_host(job, src.ctypes.data, src.strides[0], out.ctypes.data, out.strides[0], 0, oh)
    assert ok, dll.iwb200_job_last_error(job).decode()
    r0, nr, w = C.c_int(), C.c_int(), C.c_int()
    last = -1                                                      # the channel processed last, which is the one the oracle keeps
    assert dll.iwb200_job_read_intermediate(job, last, None, 0, C.byref(r0), C.byref(nr), C.byref(w))
    assert (r0.value, nr.value, w.value) == (0, oh, iw), "the whole image must be one chunk at this size"
    got = np.zeros((nr.value, w.value), np.float32)
    assert dll.iwb200_job_read_intermediate(job, last, got.ctypes.data, got.size, C.byref(r0), C.byref(nr), C.byref(w))
    dll.iwb200_job_destroy(job)
    prod.iw_destroy_context(ctx)

    assert np.array_equal(out, opx), "output bytes differ from the oracle"
    assert want.shape == got.shape
    bad = np.argwhere(want.view(np.uint32) != got.view(np.uint32))
    detail = "; ".join(f"(row {j}, col {i}): oracle {want[j, i]!r} cuda {got[j, i]!r}" for j, i in bad[:6])
    assert len(bad) == 0, f"{len(bad)} of {got.size} FP32 intermediate samples differ: {detail}"


# ---- the sign of zero ---------------------------------------------------------------------------------------------
# The uniform paths assign an output's first product instead of adding it to a zeroed accumulator (stream2.cuh:
# IWB2_FIRST_ASSIGN).  A black sample under a negative filter lobe makes that product -0.0 where the reference holds
# +0.0 + -0.0 = +0.0; the next tap with a positive weight brings the sum back to +0.0, so the FP32 planes must stay
# bit-identical -- including the sign of their zeros -- on images with black regions, rows and columns.
ZERO_CASES = [
    ("black_down2_lanczos_rgb", 512, 384, RGB, 8, 256, 192, "-filter lanczos -dither o", {}),
    ("black_down3_catrom_holes", 480, 360, RGB, 8, 160, 120, "-filter catrom -dither o", {}),
    ("black_down2_gray_as_rgb", 512, 384, GRAY, 8, 256, 192, "-filter lanczos -dither r", {}),
    ("black_rgb16_stream", 320, 200, RGB, 16, 160, 100, "-filter lanczos -dither o", {}),
    ("black_down4_mitchell_gray", 512, 384, GRAY, 8, 128, 96, "-filter mitchell -grayscale", {}),
    ("black_rgba8_down2_gather", 400, 300, L.IW_IMGTYPE_RGBA, 8, 200, 150, "-filter catrom -bkgd 246", {}),
    ("black_rgba16_up_gather", 200, 150, L.IW_IMGTYPE_RGBA, 16, 500, 380, "-filter mitchell -bkgd 369", {}),
    ("black_ga8_down_gather", 400, 300, L.IW_IMGTYPE_GRAYA, 8, 260, 200, "-filter catrom", {}),
    ("black_rgba8_down3_tile", 420, 300, L.IW_IMGTYPE_RGBA, 8, 140, 100, "-filter catrom -bkgd 246", {}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("pattern", ["all", "blocks", "stripes"])
@pytest.mark.parametrize("case", ZERO_CASES, ids=lambda c: c[0])
def test_black_regions_keep_the_sign_of_zero(prod, oracle_port, case, pattern):
    name, iw, ih, typ, depth, ow, oh, extra, opts = case
    cfg = configs.Config(name, iw, ih, typ, depth, ow, oh, f"-width {ow} -height {oh} {extra}")
    rows = configs.synth_rows(prod, cfg).copy()
    bpp = L.imgtype_num_channels(typ) * depth // 8
    if pattern == "all":
        rows[:] = 0
    elif pattern == "blocks":                    # black rectangles with noise around them (and noise islands inside)
        rows[ih // 5: ih // 2, (iw // 6) * bpp: (iw // 2) * bpp] = 0
        rows[ih // 2 + 7:, : (iw // 3) * bpp] = 0
        rows[ih // 2 + 40: ih // 2 + 43, (iw // 8) * bpp: (iw // 8 + 2) * bpp] = 0xc8
    else:                                        # single black rows / columns between noise
        rows[::7] = 0
        px = rows.reshape(ih, iw, bpp)
        px[:, ::5] = 0
    _check_planes(prod, cfg, rows, rows, opts)


# ---- the same comparison against the UNMODIFIED reference build --------------------------------------------------
# The reference does not expose its intermediate, but every allocation it makes goes through the context's
# mallocfn / freefn (src/imagew.h:467-491).  A freefn that copies the block of in_w * out_h floats just before the
# reference frees it at the end of iw_process_internal (src/imagew-main.c:1446) captures `intermediate32` as the last
# processed channel left it -- so the oracle port is not the only witness of the FP32 planes.
_MALLOCFN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t)
_FREEFN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


class _InitParams(C.Structure):
    _fields_ = [("api_version", C.c_int), ("userdata", C.c_void_p), ("mallocfn", _MALLOCFN), ("freefn", _FREEFN)]


def _reference_intermediate(ref, cfg, rows):
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.calloc.restype = C.c_void_p
    libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    sizes, snaps = {}, []
    want_bytes = cfg.in_w * cfg.out_h * 4

    def mallocfn(_ud, flags, n):
        p = libc.calloc(1, n) if flags & 1 else libc.malloc(n)
        if p:
            sizes[p] = n
        return p

    def freefn(_ud, p):
        if p and sizes.pop(p, None) == want_bytes:
            snaps.append(np.frombuffer(C.string_at(p, want_bytes), dtype=np.float32).reshape(cfg.out_h, cfg.in_w).copy())
        libc.free(p)

    params = _InitParams(0, None, _MALLOCFN(mallocfn), _FREEFN(freefn))
    ctx = ref.iw_create_context(C.cast(C.byref(params), C.c_void_p))
    assert ctx
    nch = L.imgtype_num_channels(cfg.imgtype)
    ref.set_input_from_numpy(ctx, rows[:, : cfg.in_w * nch * (cfg.bit_depth // 8)], cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    cfg.apply(ref, ctx)
    assert ref.iw_process_image(ctx), ref.errmsg(ctx)
    _, out = ref.get_output_numpy(ctx)
    out = out.copy()
    ref.iw_destroy_context(ctx)
    # (intermediate32 is the first block of that size to be freed: src/imagew-main.c:1446-1448)
    assert len(snaps) >= 1, "the reference freed no block of the intermediate's size"
    return snaps[0], out


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("down2_lanczos_rgb", "odd_catrom_gray", "up_mitchell_rgb", "gray16_stream",
                                                            "down3_catrom_holes", "rgba16_up_gather")], ids=lambda c: c[0])
def test_fp32_intermediate_is_bit_identical_to_the_reference_build(prod, ref, case):
    name, iw, ih, typ, depth, ow, oh, extra, opts = case
    cfg = configs.Config(name, iw, ih, typ, depth, ow, oh, f"-width {ow} -height {oh} {extra}")
    rows = configs.synth_rows(prod, cfg)
    want, ref_out = _reference_intermediate(ref, cfg, rows)

    dll = prod.dll
    _job_api(dll)
    ctx = prod.iw_create_context(None)
    prod.set_input_from_numpy(ctx, rows, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    cfg.apply(prod, ctx)
    job = dll.iwb200_job_create(ctx, 0)
    assert job, prod.errmsg(ctx)
    info = Info()
    dll.iwb200_job_get_info(job, C.byref(info))
    out = np.zeros((oh, info.out_bpr), np.uint8)
    src = np.ascontiguousarray(rows)
    assert dll.iwb200_job_run_host(job, src.ctypes.data, src.strides[0], out.ctypes.data, out.strides[0], 0, oh)
    r0, nr, w = C.c_int(), C.c_int(), C.c_int()
    got = np.zeros((oh, iw), np.float32)
    assert dll.iwb200_job_read_intermediate(job, -1, got.ctypes.data, got.size, C.byref(r0), C.byref(nr), C.byref(w))
    dll.iwb200_job_destroy(job)
    prod.iw_destroy_context(ctx)
    assert np.array_equal(out, ref_out)
    bad = int(np.count_nonzero(want.view(np.uint32) != got.view(np.uint32)))
    assert bad == 0, f"{bad} of {got.size} FP32 intermediate samples differ from the reference build's"


def test_reference_intermediate_snapshot_agrees_with_the_oracle_port(ref, oracle_port):
    """Host-only: the freefn snapshot of the real reference equals the plane the C restatement keeps."""
    from oracle_port import OraclePort
    cfg = configs.Config("snap", 96, 64, RGB, 8, 40, 30, "-width 40 -height 30 -filter lanczos -dither o")
    rng = np.random.default_rng(5)
    rows = rng.integers(0, 256, (cfg.in_h, cfg.in_w * 3), dtype=np.uint8)
    want, _ = _reference_intermediate(ref, cfg, rows)
    olib = OraclePort()
    octx = olib.iw_create_context(None)
    olib.set_input_from_numpy(octx, rows, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    cfg.apply(olib, octx)
    assert olib.iw_process_image(octx)
    assert np.array_equal(olib.intermediate(octx).view(np.uint32), want.view(np.uint32))
