"""iwb200_batch_*: many same-shaped images through one plan, copies and kernels of consecutive images overlapped
(csrc/pipeline.cu).  Every image of a batch of distinct noise images must equal the reference's output for that image
(BASELINE configuration 4: 1920x1080 -> 256x144 thumbnails, automatic filter and box filter), also with several slots
on one device, with padded host rows, and for a job that runs on the generic kernels."""
import ctypes as C

import numpy as np
import pytest

from imageworsener_b200 import configs, libimagew as L


def _api(dll):
    dll.iwb200_batch_create.restype = C.c_void_p
    dll.iwb200_batch_create.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int]
    dll.iwb200_batch_submit.restype = C.c_int
    dll.iwb200_batch_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    dll.iwb200_batch_wait.restype = C.c_int
    dll.iwb200_batch_wait.argtypes = [C.c_void_p]
    dll.iwb200_batch_destroy.argtypes = [C.c_void_p]
    dll.iwb200_batch_last_error.restype = C.c_char_p
    dll.iwb200_batch_last_error.argtypes = [C.c_void_p]


def run_batch(lib, cfg, images, devices=(0,), depth=3, pad=0):
    """images: list of (H, bpr) uint8 arrays.  Returns the list of outputs."""
    dll = lib.dll
    _api(dll)
    ctx = lib.iw_create_context(None)
    img = L.iw_image()
    img.imgtype, img.bit_depth, img.sampletype = cfg.imgtype, cfg.bit_depth, 0
    img.width, img.height, img.bpr, img.pixels = cfg.in_w, cfg.in_h, cfg.in_bpr, None
    lib.iw_set_input_image(ctx, C.byref(img))
    cfg.apply(lib, ctx)
    devs = (C.c_int * len(devices))(*devices)
    b = dll.iwb200_batch_create(ctx, devs, len(devices), depth)
    assert b, lib.errmsg(ctx)
    from test_intermediate import Info
    dll.iwb200_batch_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
    info = Info()
    assert dll.iwb200_batch_get_info(b, C.byref(info))
    out_bpr = info.out_bpr
    outs = [np.zeros((cfg.out_h, out_bpr), np.uint8) for _ in images]
    srcs = []
    for im in images:
        if pad:
            p = np.full((im.shape[0], im.shape[1] + pad), 0x5A, np.uint8)
            p[:, : im.shape[1]] = im
            im = p
        srcs.append(np.ascontiguousarray(im))
    try:
        for s, o in zip(srcs, outs):
            assert dll.iwb200_batch_submit(b, s.ctypes.data, s.shape[1], o.ctypes.data, o.shape[1]), dll.iwb200_batch_last_error(b)
        assert dll.iwb200_batch_wait(b), dll.iwb200_batch_last_error(b)
    finally:
        dll.iwb200_batch_destroy(b)
        lib.iw_destroy_context(ctx)
    return outs


def _images(lib, cfg, n):
    return [configs.synth_rows(lib, cfg, seed=configs.XORSHIFT_SEED + 17 * k) for k in range(n)]


@pytest.mark.parametrize("name", ["cfg4auto", "cfg4box"])
def test_host_logic_batch_matches_reference(emu, ref, name):
    cfg = configs.get(name, scale=4)
    imgs = _images(emu, cfg, 5)
    outs = run_batch(emu, cfg, imgs, depth=2)
    for im, o in zip(imgs, outs):
        _, want = configs.run_host(ref, cfg, im)
        assert np.array_equal(o, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name,depth,pad", [("cfg4auto", 3, 0), ("cfg4box", 1, 0), ("cfg4auto", 4, 16), ("cfg4box", 2, 7)])
def test_cuda_batch_matches_reference(prod, ref, name, depth, pad):
    cfg = configs.get(name)
    imgs = _images(prod, cfg, 12)
    outs = run_batch(prod, cfg, imgs, depth=depth, pad=pad)
    for k, (im, o) in enumerate(zip(imgs, outs)):
        _, want = configs.run_host(ref, cfg, im)
        assert np.array_equal(o, want), f"image {k}"


@pytest.mark.gpu
def test_cuda_batch_on_generic_kernels_and_two_slots_sets(prod, ref):
    """A job the streaming kernels do not take (RGBA16 with a replicated edge): the batch runs the generic kernels."""
    cfg = configs.Config("gen", 200, 160, L.IW_IMGTYPE_RGBA, 16, 150, 120, "-width 150 -height 120 -filter catrom -edge r -depth 16 -bkgd 369")
    imgs = _images(prod, cfg, 7)
    outs = run_batch(prod, cfg, imgs, devices=(0, 0), depth=2)
    for im, o in zip(imgs, outs):
        _, want = configs.run_host(ref, cfg, im)
        assert np.array_equal(o, want)
