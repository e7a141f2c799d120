"""Python shim that lets the job/harness code drive oracle/libiworacle.so (the plain-C restatement) through the
same method names as a libimagew binding.  The shim only records settings into the oracle's flat parameter
struct; all pixel maths happens in C (oracle/iw_oracle.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from imageworsener_b200 import libimagew as L
import oracle  # noqa: E402  (test infrastructure: the reference build / port)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Params(C.Structure):
    _fields_ = [
        ("imgtype", C.c_int), ("bit_depth", C.c_int), ("sampletype", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("bpr", C.c_size_t), ("pixels", C.c_void_p), ("orient", C.c_uint), ("in_maxcode", C.c_int * 4),
        ("cs_in_type", C.c_int), ("cs_in_gamma", C.c_double), ("in_label_set", C.c_int), ("in_label", C.c_double * 4),
        ("profile", C.c_uint), ("out_depth_req", C.c_int), ("out_maxcode_req", C.c_int * 5),
        ("color_count_req", C.c_int * 5), ("dither_family", C.c_int * 5), ("dither_subtype", C.c_int * 5),
        ("canvas_w", C.c_int), ("canvas_h", C.c_int), ("out_true_valid", C.c_int), ("out_true_w", C.c_double),
        ("out_true_h", C.c_double), ("crop_x", C.c_int), ("crop_y", C.c_int), ("crop_w", C.c_int), ("crop_h", C.c_int),
        ("cs_out_valid", C.c_int), ("cs_out_type", C.c_int), ("cs_out_gamma", C.c_double),
        ("family", C.c_int * 2), ("blur", C.c_double * 2), ("param1", C.c_double * 2), ("param2", C.c_double * 2),
        ("edge", C.c_int * 2), ("translate", C.c_double * 2), ("chan_offset", (C.c_double * 3) * 2),
        ("to_grayscale", C.c_int), ("gsf", C.c_int), ("gray_w", C.c_double * 3),
        ("no_gamma", C.c_int), ("intclamp", C.c_int), ("negate", C.c_int),
        ("bkgd_valid", C.c_int), ("bkgd", C.c_double * 4), ("checker_req", C.c_int), ("bkgd2", C.c_double * 4),
        ("check_size", C.c_int), ("check_ox", C.c_int), ("check_oy", C.c_int),
        ("use_label_from_file", C.c_int), ("suppress_out_label", C.c_int), ("out_label_valid", C.c_int),
        ("out_label", C.c_double * 4), ("random_seed", C.c_int), ("max_width", C.c_int), ("max_height", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [
        ("imgtype", C.c_int), ("bit_depth", C.c_int), ("sampletype", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("bpr", C.c_size_t), ("pixels", C.c_void_p), ("cs_out_type", C.c_int), ("cs_out_gamma", C.c_double),
        ("reduced_maxcolors", C.c_int), ("maxcolorcode", C.c_uint * 5), ("has_bkgdlabel", C.c_int),
        ("bkgdlabel", C.c_double * 4), ("last_intermediate", C.c_void_p), ("inter_w", C.c_int), ("inter_h", C.c_int),
    ]


class _Ctx:
    def __init__(self, dll):
        self.p = Params()
        dll.iwo_default_params(C.byref(self.p))
        self.res = None
        self.err = ""
        self.keep = None
        self.randomize = 0


def _obj(x):
    return x._obj if hasattr(x, "_obj") else x


class OraclePort:
    is_product = False

    def __init__(self, so=None):
        so = so or os.path.join(ROOT, "oracle", "libiworacle.so")
        self.dll = C.CDLL(so, mode=os.RTLD_LOCAL)
        self.dll.iwo_process.argtypes = [C.POINTER(Params), C.POINTER(Result), C.c_char_p, C.c_int]
        self.dll.iwo_to_linear.restype = C.c_double
        self.dll.iwo_to_linear.argtypes = [C.c_double, C.c_int, C.c_double]
        self.dll.iwo_prng_nth.restype = C.c_uint32
        self.dll.iwo_prng_nth.argtypes = [C.c_int, C.c_uint64]

    # life cycle
    def iw_create_context(self, _):
        return _Ctx(self.dll)

    def iw_destroy_context(self, ctx):
        if ctx.res is not None:
            self.dll.iwo_free_result(C.byref(ctx.res))
            ctx.res = None

    def iw_set_warning_fn(self, ctx, fn):
        pass

    def iw_set_max_malloc(self, ctx, n):
        pass

    def iw_set_option(self, ctx, k, v):
        pass

    def iw_set_allow_opt(self, ctx, opt, n):
        pass

    def errmsg(self, ctx):
        return ctx.err

    # input
    def set_input_from_numpy(self, ctx, arr, imgtype, bit_depth, width, height, sampletype=0, native_grayscale=0):
        ctx.keep = np.ascontiguousarray(arr, dtype=np.uint8)
        p = ctx.p
        p.imgtype, p.bit_depth, p.sampletype, p.width, p.height = imgtype, bit_depth, sampletype, width, height
        p.bpr = ctx.keep.shape[1]
        p.pixels = ctx.keep.ctypes.data
        ctx.native_grayscale = native_grayscale

    def iw_set_input_colorspace(self, ctx, cs):
        cs = _obj(cs)
        t, g = cs.cstype, cs.gamma
        if t == L.IW_CSTYPE_GAMMA and 0.999995 <= g <= 1.000005:
            t = L.IW_CSTYPE_LINEAR
        ctx.p.cs_in_type, ctx.p.cs_in_gamma = t, g

    def iw_set_output_colorspace(self, ctx, cs):
        cs = _obj(cs)
        t, g = cs.cstype, cs.gamma
        if t == L.IW_CSTYPE_GAMMA and 0.999995 <= g <= 1.000005:
            t = L.IW_CSTYPE_LINEAR
        ctx.p.cs_out_valid, ctx.p.cs_out_type, ctx.p.cs_out_gamma = 1, t, g

    def iw_make_linear_csdescr(self, cs):
        cs = _obj(cs); cs.cstype, cs.gamma, cs.srgb_intent = L.IW_CSTYPE_LINEAR, 0.0, 0

    def iw_make_srgb_csdescr_2(self, cs):
        cs = _obj(cs); cs.cstype, cs.gamma = L.IW_CSTYPE_SRGB, 0.0

    def iw_make_rec709_csdescr(self, cs):
        cs = _obj(cs); cs.cstype, cs.gamma = L.IW_CSTYPE_REC709, 0.0

    def iw_make_gamma_csdescr(self, cs, g):
        cs = _obj(cs)
        cs.cstype, cs.gamma, cs.srgb_intent = L.IW_CSTYPE_GAMMA, min(max(g, 0.1), 10.0), 0
        if 0.999995 <= cs.gamma <= 1.000005:
            cs.cstype = L.IW_CSTYPE_LINEAR

    def iw_convert_sample_to_linear(self, v, cs):
        cs = _obj(cs)
        return self.dll.iwo_to_linear(v, cs.cstype, cs.gamma)

    def iw_set_input_bkgd_label(self, ctx, r, g, b):
        ctx.p.in_label_set = 1
        ctx.p.in_label[:] = [r, g, b, 1.0]

    def iw_set_input_max_color_code(self, ctx, ch, c):
        if 0 <= ch < 4:
            ctx.p.in_maxcode[ch] = c

    def iw_reorient_image(self, ctx, x):
        tbl = [4, 6, 5, 7, 0, 2, 1, 3]
        p = ctx.p
        x &= 7
        if x & 4:
            p.orient = tbl[p.orient]
            p.width, p.height = p.height, p.width
        p.orient ^= (x & 3)

    # settings
    def iw_set_random_seed(self, ctx, randomize, seed):
        ctx.p.random_seed = seed

    def iw_set_value(self, ctx, code, n):
        p = ctx.p
        if code == L.IW_VAL_CVT_TO_GRAYSCALE: p.to_grayscale = n
        elif code == L.IW_VAL_DISABLE_GAMMA: p.no_gamma = n
        elif code == L.IW_VAL_INT_CLAMP: p.intclamp = n
        elif code == L.IW_VAL_EDGE_POLICY_X: p.edge[0] = n
        elif code == L.IW_VAL_EDGE_POLICY_Y: p.edge[1] = n
        elif code == L.IW_VAL_GRAYSCALE_FORMULA: p.gsf = n
        elif code == L.IW_VAL_USE_BKGD_LABEL: p.use_label_from_file = n
        elif code == L.IW_VAL_NO_BKGD_LABEL: p.suppress_out_label = n
        elif code == L.IW_VAL_NEGATE_TARGET: p.negate = n
        elif code == L.IW_VAL_MAX_WIDTH: p.max_width = n
        elif code == L.IW_VAL_MAX_HEIGHT: p.max_height = n

    def iw_get_value(self, ctx, code):
        if code == L.IW_VAL_INPUT_WIDTH: return max(1, ctx.p.width)
        if code == L.IW_VAL_INPUT_HEIGHT: return max(1, ctx.p.height)
        if code == L.IW_VAL_INPUT_NATIVE_GRAYSCALE: return ctx.native_grayscale
        return 0

    def iw_set_value_dbl(self, ctx, code, v):
        if code == L.IW_VAL_TRANSLATE_X: ctx.p.translate[0] = v
        elif code == L.IW_VAL_TRANSLATE_Y: ctx.p.translate[1] = v

    def iw_set_grayscale_weights(self, ctx, r, g, b):
        tot = r + g + b
        if tot == 0.0:
            tot = 1.0
        ctx.p.gray_w[:] = [r / tot, g / tot, b / tot]

    def iw_get_profile_by_fmt(self, fmt):
        return oracle.load_reference().iw_get_profile_by_fmt(fmt) if os.path.exists(oracle.REFERENCE_SO) else _PROFILES[fmt]

    def iw_set_output_profile(self, ctx, n): ctx.p.profile = n
    def iw_set_output_depth(self, ctx, bps): ctx.p.out_depth_req = bps

    def iw_set_output_max_color_code(self, ctx, ct, n):
        if 0 <= ct < 5:
            ctx.p.out_maxcode_req[ct] = n

    @staticmethod
    def _fan(ct):
        ks = [ct] if 0 <= ct < 5 else []
        if ct == L.IW_CHANNELTYPE_ALL:
            ks.append(3)
        if ct in (L.IW_CHANNELTYPE_ALL, L.IW_CHANNELTYPE_NONALPHA):
            ks += [0, 1, 2, 4]
        return ks

    def iw_set_dither_type(self, ctx, ct, f, s):
        for k in self._fan(ct):
            ctx.p.dither_family[k], ctx.p.dither_subtype[k] = f, s

    def iw_set_color_count(self, ctx, ct, c):
        for k in self._fan(ct):
            ctx.p.color_count_req[k] = c

    def iw_set_channel_offset(self, ctx, ct, dim, offs):
        if 0 <= ct <= 2:
            ctx.p.chan_offset[dim if dim in (0, 1) else 0][ct] = offs

    def iw_set_apply_bkgd_2(self, ctx, clr):
        ctx.p.bkgd_valid = 1
        ctx.p.bkgd[:] = list(_obj(clr).c)

    def iw_set_bkgd_checkerboard_2(self, ctx, size, clr):
        ctx.p.checker_req, ctx.p.check_size = 1, size
        ctx.p.bkgd2[:] = list(_obj(clr).c)

    def iw_set_bkgd_checkerboard_origin(self, ctx, x, y):
        ctx.p.check_ox, ctx.p.check_oy = x, y

    def iw_set_output_bkgd_label_2(self, ctx, clr):
        ctx.p.out_label_valid = 1
        ctx.p.out_label[:] = list(_obj(clr).c)

    def iw_set_resize_alg(self, ctx, dim, family, blur, p1, p2):
        d = dim if dim in (0, 1) else 0
        ctx.p.family[d], ctx.p.blur[d], ctx.p.param1[d], ctx.p.param2[d] = family, blur, p1, p2

    def iw_set_output_canvas_size(self, ctx, w, h): ctx.p.canvas_w, ctx.p.canvas_h = w, h

    def iw_set_output_image_size(self, ctx, w, h):
        ctx.p.out_true_valid, ctx.p.out_true_w, ctx.p.out_true_h = 1, max(w, 0.01), max(h, 0.01)

    def iw_set_input_crop(self, ctx, x, y, w, h):
        ctx.p.crop_x, ctx.p.crop_y, ctx.p.crop_w, ctx.p.crop_h = x, y, w, h

    # run
    def iw_process_image(self, ctx):
        ctx.res = Result()
        buf = C.create_string_buffer(256)
        ok = self.dll.iwo_process(C.byref(ctx.p), C.byref(ctx.res), buf, 256)
        ctx.err = buf.value.decode("utf-8", "replace")
        return ok

    def get_output_numpy(self, ctx):
        r = ctx.res
        img = L.iw_image()
        img.imgtype, img.bit_depth, img.sampletype = r.imgtype, r.bit_depth, r.sampletype
        img.width, img.height, img.bpr = r.width, r.height, r.bpr
        img.reduced_maxcolors = r.reduced_maxcolors
        for k in range(5):
            img.maxcolorcode[k] = r.maxcolorcode[k]
        img.has_bkgdlabel = r.has_bkgdlabel
        for k in range(4):
            img.bkgdlabel.c[k] = r.bkgdlabel[k]
        n = r.bpr * r.height
        px = np.frombuffer((C.c_uint8 * n).from_address(r.pixels), np.uint8).reshape(r.height, r.bpr).copy()
        return img, px

    def iw_get_output_colorspace(self, ctx, cs):
        cs = _obj(cs)
        cs.cstype, cs.gamma = ctx.res.cs_out_type, ctx.res.cs_out_gamma

    def intermediate(self, ctx):
        r = ctx.res
        n = r.inter_w * r.inter_h
        return np.frombuffer((C.c_float * n).from_address(r.last_intermediate), np.float32).reshape(r.inter_h, r.inter_w).copy()


_PROFILES = {1: 0x19f7f, 2: 0x2, 3: 0x9a00, 4: 0x186b, 5: 0x46003, 6: 0x1, 8: 0x804a, 9: 0x8, 10: 0x8042, 11: 0x8040,
             12: 0x804b}
