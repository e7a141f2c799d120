"""The fused vertical + horizontal kernel (generation 3, csrc/stream3*.cu): one persistent, warp-specialised kernel
whose V warps and H warps exchange the FP32 intermediate through L2 under release / acquire counters.  Selected with the
context option cuda:gen=3 (it is opt-in: on the headline configuration it measured slower than the two separate
kernels) on shapes that cover: 6 and
4 open windows at step 2 (Lanczos3 / cubic 2x reductions, the pairs the kernel is built for), RGB and gray, every fast
epilogue (none / nearest-colour table, ordered, halftone, random), ragged widths and heights (partial warp columns,
partial last group of 32 rows), output-row bands, different V : H warp splits and segment lengths, repeated launches
of one plan (the counters are generation-based, never reset), and the FP32 intermediate itself.

All tests call through the C ABI and compare with the unmodified reference build, byte for byte.
"""
import ctypes as C

import numpy as np
import pytest

from imageworsener_b200 import configs, libimagew as L

RGB, GRAY = L.IW_IMGTYPE_RGB, L.IW_IMGTYPE_GRAY

SHAPES = [
    ("lanczos_o", 1024, 768, RGB, 512, 384, "-filter lanczos -dither o"),
    ("lanczos_none", 1024, 768, RGB, 512, 384, "-filter lanczos"),
    ("lanczos_halftone_gray", 1024, 768, GRAY, 512, 384, "-filter lanczos -grayscale -dither halftone"),
    ("lanczos_r", 1280, 400, RGB, 640, 200, "-filter lanczos -dither r"),
    ("lanczos_r2_gray", 1280, 400, GRAY, 640, 200, "-filter lanczos -grayscale -dither r2"),
    ("catrom_r", 1024, 768, RGB, 512, 384, "-filter catrom -dither r"),
    ("mitchell_gray", 1024, 768, GRAY, 512, 384, "-filter mitchell -grayscale"),
    ("catrom_o", 2048, 512, RGB, 1024, 256, "-filter catrom -dither o"),
    ("ragged", 1002, 702, RGB, 501, 351, "-filter lanczos -dither o"),             # partial warp column, partial group
    ("ragged_gray", 1238, 334, GRAY, 619, 167, "-filter catrom -grayscale -dither o"),
    ("tall", 520, 4000, RGB, 260, 2000, "-filter lanczos -dither o"),              # many groups, few columns
    ("wide", 8000, 200, RGB, 4000, 100, "-filter lanczos -dither halftone"),       # many segments, 4 groups
    ("big", 3840, 2160, RGB, 1920, 1080, "-filter lanczos -dither o"),
]


def _cfg(shape):
    name, iw, ih, typ, ow, oh, extra = shape
    return configs.Config(name, iw, ih, typ, 8, ow, oh, f"-width {ow} -height {oh} {extra}")


def _job_api(dll):
    dll.iwb200_job_create.restype = C.c_void_p
    dll.iwb200_job_create.argtypes = [C.c_void_p, C.c_int]
    dll.iwb200_job_destroy.argtypes = [C.c_void_p]
    dll.iwb200_job_last_kernel_generation.restype = C.c_int
    dll.iwb200_job_last_kernel_generation.argtypes = [C.c_void_p]
    dll.iwb200_job_run_host.restype = C.c_int
    dll.iwb200_job_run_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    dll.iwb200_job_last_error.restype = C.c_char_p
    dll.iwb200_job_last_error.argtypes = [C.c_void_p]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
def test_fused_kernel_matches_reference(prod, ref, shape):
    cfg = _cfg(shape)
    rows = configs.synth_rows(prod, cfg)
    ia, a = configs.run_host(prod, cfg, rows, options={"cuda:gen": 3})
    ib, b = configs.run_host(ref, cfg, rows)
    assert (ia.width, ia.height, ia.bit_depth, ia.imgtype, ia.bpr) == (ib.width, ib.height, ib.bit_depth, ib.imgtype, ib.bpr)
    nd = int(np.count_nonzero(a != b))
    assert nd == 0, f"{nd} differing bytes of {a.size}"


@pytest.mark.gpu
def test_fused_kernel_is_really_taken_and_relaunchable(prod, ref):
    """The job reports generation 3 for its last launch, and ten launches of the same plan (the cross-CTA counters are
    never reset, launch n waits for n times the announcements) all give the reference's bytes."""
    cfg = _cfg(SHAPES[0])
    rows = configs.synth_rows(prod, cfg)
    _, b = configs.run_host(ref, cfg, rows)
    dll = prod.dll
    _job_api(dll)
    ctx = prod.iw_create_context(None)
    prod.set_input_from_numpy(ctx, rows, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    prod.iw_set_option(ctx, b"cuda:gen", b"3")
    cfg.apply(prod, ctx)
    job = dll.iwb200_job_create(ctx, 0)
    assert job, prod.errmsg(ctx)
    try:
        out = np.zeros_like(b)
        for k in range(10):
            out[:] = 0
            ok = dll.iwb200_job_run_host(job, rows.ctypes.data, rows.shape[1], out.ctypes.data, out.shape[1], 0, cfg.out_h)
            assert ok, dll.iwb200_job_last_error(job)
            assert dll.iwb200_job_last_kernel_generation(job) == 3
            assert np.array_equal(out, b), f"launch {k}"
    finally:
        dll.iwb200_job_destroy(job)
        prod.iw_destroy_context(ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("nv,seg", [(4, 32), (8, 64), (12, 200), (15, 128), (1, 512)])
def test_fused_kernel_warp_split_and_segments_do_not_matter(prod, ref, nv, seg):
    cfg = _cfg(SHAPES[8])
    rows = configs.synth_rows(prod, cfg)
    _, a = configs.run_host(prod, cfg, rows, options={"cuda:gen": 3, "cuda:nv": nv, "cuda:seg3": seg})
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[5], SHAPES[8]], ids=lambda s: s[0])
@pytest.mark.parametrize("nbands", [2, 3])
def test_fused_kernel_bands_concatenate_to_reference(prod, ref, shape, nbands):
    cfg = _cfg(shape)
    rows = configs.synth_rows(prod, cfg)
    parts = [configs.run_host(prod, cfg, rows, options={"cuda:gen": 3, "cuda:band": f"{i}/{nbands}"})[1] for i in range(nbands)]
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(np.concatenate(parts, axis=0), b)
