"""ctypes binding of the libimagew C API surface that the pixel pipeline is driven through.

The same binding class loads either implementation of that API:

  * ``imageworsener_b200/libimagew_b200.so``  -- this repo's CUDA implementation (the product)
  * ``oracle/_ref/libiwref.so``               -- the unmodified reference, test infrastructure only

so parity tests issue literally the same calls against both.  Names, argument meaning and constants
follow the reference header (reference: src/imagew.h:40-410 constants and structs, :497-703 functions).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

# ---- constants (reference: src/imagew.h:51-335) -------------------------------------------------
IW_VAL_CVT_TO_GRAYSCALE = 11
IW_VAL_DISABLE_GAMMA = 13
IW_VAL_NO_CSLABEL = 14
IW_VAL_INT_CLAMP = 15
IW_VAL_GRAYSCALE_FORMULA = 16
IW_VAL_EDGE_POLICY_X = 17
IW_VAL_EDGE_POLICY_Y = 18
IW_VAL_INPUT_NATIVE_GRAYSCALE = 20
IW_VAL_INPUT_WIDTH = 21
IW_VAL_INPUT_HEIGHT = 22
IW_VAL_USE_BKGD_LABEL = 37
IW_VAL_TRANSLATE_X = 42
IW_VAL_TRANSLATE_Y = 43
IW_VAL_MAX_WIDTH = 45
IW_VAL_MAX_HEIGHT = 46
IW_VAL_NO_BKGD_LABEL = 48
IW_VAL_NEGATE_TARGET = 53

IW_FORMAT_PNG, IW_FORMAT_JPEG, IW_FORMAT_BMP, IW_FORMAT_TIFF, IW_FORMAT_MIFF = 1, 2, 3, 4, 5
IW_FORMAT_PNM, IW_FORMAT_PBM, IW_FORMAT_PGM, IW_FORMAT_PPM, IW_FORMAT_PAM = 8, 9, 10, 11, 12

IW_PROFILE_TRANSPARENCY = 0x0001
IW_PROFILE_GRAYSCALE = 0x0002
IW_PROFILE_16BPS = 0x0040
IW_PROFILE_ALWAYSLINEAR = 0x2000
IW_PROFILE_HDRI = 0x4000
IW_PROFILE_REDUCEDBITDEPTHS = 0x8000
IW_PROFILE_PNG_BKGD = 0x10000
IW_PROFILE_RGB8_BKGD = 0x20000
IW_PROFILE_RGB16_BKGD = 0x40000

IW_RESIZETYPE_AUTO = 0x01
IW_RESIZETYPE_NULL = 0x02
IW_RESIZETYPE_NEAREST = 0x10
IW_RESIZETYPE_MIX = 0x11
IW_RESIZETYPE_BOX = 0x20
IW_RESIZETYPE_TRIANGLE = 0x21
IW_RESIZETYPE_QUADRATIC = 0x22
IW_RESIZETYPE_GAUSSIAN = 0x23
IW_RESIZETYPE_HERMITE = 0x24
IW_RESIZETYPE_CUBIC = 0x25
IW_RESIZETYPE_BOXAVG = 0x26
IW_RESIZETYPE_SINC = 0x30
IW_RESIZETYPE_LANCZOS = 0x31
IW_RESIZETYPE_HANNING = 0x32
IW_RESIZETYPE_BLACKMAN = 0x33

IW_IMGTYPE_GRAY, IW_IMGTYPE_GRAYA, IW_IMGTYPE_RGB, IW_IMGTYPE_RGBA = 0x0001, 0x0101, 0x0010, 0x0110
IW_SAMPLETYPE_UINT, IW_SAMPLETYPE_FLOATINGPOINT = 0, 1

IW_CHANNELTYPE_RED, IW_CHANNELTYPE_GREEN, IW_CHANNELTYPE_BLUE = 0, 1, 2
IW_CHANNELTYPE_ALPHA, IW_CHANNELTYPE_GRAY = 3, 4
IW_CHANNELTYPE_ALL, IW_CHANNELTYPE_NONALPHA = 10, 11

IW_CSTYPE_SRGB, IW_CSTYPE_LINEAR, IW_CSTYPE_GAMMA, IW_CSTYPE_REC709 = 0, 1, 2, 3
IW_DIMENSION_H, IW_DIMENSION_V = 0, 1

IW_DITHERFAMILY_NONE, IW_DITHERFAMILY_ORDERED, IW_DITHERFAMILY_ERRDIFF, IW_DITHERFAMILY_RANDOM = 0, 1, 2, 3
IW_DITHERSUBTYPE_HALFTONE = 1
IW_DITHERSUBTYPE_SAMEPATTERN = 1

IW_GSF_STANDARD, IW_GSF_COMPATIBLE, IW_GSF_WEIGHTED, IW_GSF_ORDERBYVALUE = 0, 1, 2, 3
IW_EDGE_POLICY_REPLICATE, IW_EDGE_POLICY_STANDARD, IW_EDGE_POLICY_TRANSPARENT = 1, 2, 3

IW_OPT_GRAYSCALE, IW_OPT_PALETTE, IW_OPT_16_TO_8, IW_OPT_STRIP_ALPHA, IW_OPT_BINARY_TRNS = 1, 2, 3, 4, 5


def imgtype_num_channels(t: int) -> int:
    """reference: src/imagew-api.c:165-176"""
    return {IW_IMGTYPE_RGBA: 4, IW_IMGTYPE_RGB: 3, IW_IMGTYPE_GRAYA: 2}.get(t, 1)


# ---- structs (reference: src/imagew.h:364-410, :482-491) -----------------------------------------
class iw_csdescr(C.Structure):
    _fields_ = [("cstype", C.c_int), ("srgb_intent", C.c_int), ("gamma", C.c_double)]


class iw_color(C.Structure):
    _fields_ = [("c", C.c_double * 4)]


class iw_image(C.Structure):
    _fields_ = [
        ("imgtype", C.c_int), ("bit_depth", C.c_int), ("sampletype", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("pixels", C.c_void_p), ("bpr", C.c_size_t),
        ("orient_transform", C.c_uint),
        ("native_grayscale", C.c_int), ("density_code", C.c_int),
        ("density_x", C.c_double), ("density_y", C.c_double),
        ("has_colorkey_trns", C.c_int), ("colorkey", C.c_uint * 3),
        ("reduced_maxcolors", C.c_int), ("maxcolorcode", C.c_uint * 5),
        ("has_bkgdlabel", C.c_int), ("bkgdlabel", iw_color),
        ("rendering_intent", C.c_int),
    ]


WARNFN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)

_P = C.c_void_p
_SIGS = {
    # name: (restype, argtypes)
    "iw_create_context": (_P, [_P]),
    "iw_destroy_context": (None, [_P]),
    "iw_process_image": (C.c_int, [_P]),
    "iw_reorient_image": (None, [_P, C.c_uint]),
    "iw_get_errorflag": (C.c_int, [_P]),
    "iw_get_errormsg": (C.c_char_p, [_P, C.c_char_p, C.c_int]),
    "iw_set_warning_fn": (None, [_P, WARNFN]),
    "iw_set_max_malloc": (None, [_P, C.c_size_t]),
    "iw_set_output_canvas_size": (None, [_P, C.c_int, C.c_int]),
    "iw_set_output_image_size": (None, [_P, C.c_double, C.c_double]),
    "iw_set_input_crop": (None, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "iw_set_output_profile": (None, [_P, C.c_uint]),
    "iw_set_output_depth": (None, [_P, C.c_int]),
    "iw_set_output_max_color_code": (None, [_P, C.c_int, C.c_int]),
    "iw_set_input_max_color_code": (None, [_P, C.c_int, C.c_int]),
    "iw_set_dither_type": (None, [_P, C.c_int, C.c_int, C.c_int]),
    "iw_set_color_count": (None, [_P, C.c_int, C.c_int]),
    "iw_set_resize_alg": (None, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "iw_set_channel_offset": (None, [_P, C.c_int, C.c_int, C.c_double]),
    "iw_set_grayscale_weights": (None, [_P, C.c_double, C.c_double, C.c_double]),
    "iw_set_input_bkgd_label": (None, [_P, C.c_double, C.c_double, C.c_double]),
    "iw_set_input_bkgd_label_2": (None, [_P, C.POINTER(iw_color)]),
    "iw_set_output_bkgd_label_2": (None, [_P, C.POINTER(iw_color)]),
    "iw_set_apply_bkgd_2": (None, [_P, C.POINTER(iw_color)]),
    "iw_set_bkgd_checkerboard_2": (None, [_P, C.c_int, C.POINTER(iw_color)]),
    "iw_set_bkgd_checkerboard_origin": (None, [_P, C.c_int, C.c_int]),
    "iw_set_output_colorspace": (None, [_P, C.POINTER(iw_csdescr)]),
    "iw_set_input_colorspace": (None, [_P, C.POINTER(iw_csdescr)]),
    "iw_make_linear_csdescr": (None, [C.POINTER(iw_csdescr)]),
    "iw_make_srgb_csdescr_2": (None, [C.POINTER(iw_csdescr)]),
    "iw_make_rec709_csdescr": (None, [C.POINTER(iw_csdescr)]),
    "iw_make_gamma_csdescr": (None, [C.POINTER(iw_csdescr), C.c_double]),
    "iw_set_random_seed": (None, [_P, C.c_int, C.c_int]),
    "iw_set_allow_opt": (None, [_P, C.c_int, C.c_int]),
    "iw_set_input_image": (None, [_P, C.POINTER(iw_image)]),
    "iw_get_output_image": (None, [_P, C.POINTER(iw_image)]),
    "iw_get_output_colorspace": (None, [_P, C.POINTER(iw_csdescr)]),
    "iw_set_value": (None, [_P, C.c_int, C.c_int]),
    "iw_get_value": (C.c_int, [_P, C.c_int]),
    "iw_set_value_dbl": (None, [_P, C.c_int, C.c_double]),
    "iw_get_value_dbl": (C.c_double, [_P, C.c_int]),
    "iw_set_option": (None, [_P, C.c_char_p, C.c_char_p]),
    "iw_get_option": (C.c_char_p, [_P, C.c_char_p]),
    "iw_convert_sample_to_linear": (C.c_double, [C.c_double, C.POINTER(iw_csdescr)]),
    "iw_convert_sample_from_linear": (C.c_double, [C.c_double, C.POINTER(iw_csdescr)]),
    "iw_get_version_int": (C.c_int, []),
    "iw_get_profile_by_fmt": (C.c_uint, [C.c_int]),
    "iw_malloc_large": (_P, [_P, C.c_size_t, C.c_size_t]),
    "iw_malloc": (_P, [_P, C.c_size_t]),
    "iw_free": (None, [_P, _P]),
    "iw_imgtype_num_channels": (C.c_int, [C.c_int]),
    "iw_calc_bytesperrow": (C.c_size_t, [C.c_int, C.c_int]),
}

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
PRODUCT_SO = os.path.join(_HERE, "libimagew_b200.so")


class IwError(RuntimeError):
    pass


class Lib:
    """One loaded implementation of the libimagew API (RTLD_LOCAL, so two can coexist)."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.dll = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)
        self.is_product = hasattr(self.dll, "iwb200_version")

    # -- convenience helpers (not part of the C API) --
    def errmsg(self, ctx) -> str:
        buf = C.create_string_buffer(256)
        self.iw_get_errormsg(ctx, buf, 256)
        return buf.value.decode("utf-8", "replace")

    def set_input_from_numpy(self, ctx, arr: np.ndarray, imgtype: int, bit_depth: int,
                             width: int, height: int, sampletype: int = IW_SAMPLETYPE_UINT,
                             native_grayscale: int = 0) -> None:
        """Hand a row-major byte buffer (H x bpr uint8) to the context.  The buffer is copied into
        iw_malloc_large memory because the context takes ownership (reference: src/imagew.h:634-637)."""
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        assert arr.ndim == 2
        nrows, bpr = arr.shape
        mem = self.iw_malloc_large(ctx, bpr, nrows)
        if not mem:
            raise IwError("iw_malloc_large failed: " + self.errmsg(ctx))
        C.memmove(mem, arr.ctypes.data, arr.nbytes)
        img = iw_image()
        img.imgtype = imgtype
        img.bit_depth = bit_depth
        img.sampletype = sampletype
        img.width = width
        img.height = height
        img.pixels = mem
        img.bpr = bpr
        img.native_grayscale = native_grayscale
        self.iw_set_input_image(ctx, C.byref(img))

    def get_output_numpy(self, ctx):
        """Returns (iw_image struct, H x bpr uint8 copy of the output pixels)."""
        img = iw_image()
        self.iw_get_output_image(ctx, C.byref(img))
        n = img.bpr * img.height
        buf = (C.c_uint8 * n).from_address(img.pixels)
        out = np.frombuffer(buf, dtype=np.uint8).reshape(img.height, img.bpr).copy()
        return img, out


_cache: dict[str, Lib] = {}


def load_product() -> Lib:
    """The CUDA implementation.  Fails loudly if the extension has not been built."""
    so = os.environ.get("IWB200_LIB") or PRODUCT_SO     # tuning: another build of the same library
    if so != PRODUCT_SO:
        if so not in _cache:
            _cache[so] = Lib(so)
        return _cache[so]
    if PRODUCT_SO not in _cache:
        if not os.path.exists(PRODUCT_SO):
            raise FileNotFoundError(
                f"{PRODUCT_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _cache[PRODUCT_SO] = Lib(PRODUCT_SO)
    return _cache[PRODUCT_SO]
