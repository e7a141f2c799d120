"""Command-line style job descriptions -> libimagew API calls.

The reference's regression suite (reference: tests/runtest) is a list of `imagew` command lines.  To let
parity tests read like those lines, this module restates the CLI's option -> API mapping
(reference: src/imagew-cmd.c:1046-1479 `iwcmd_run`, :2106-2203 option table, :2443-2740 option
arguments) for the options that influence the pixel pipeline.  Codec-only options (compression levels,
interlace, jpeg/webp knobs, density) are accepted and ignored.

Usage:
    job = Job.parse("-width 35 -height 35 -filter catrom -cc 3 -dither o", outfmt="png")
    ctx = lib.iw_create_context(None); lib.set_input_from_numpy(ctx, ...); job.apply(lib, ctx)
    lib.iw_process_image(ctx)
"""
from __future__ import annotations

import ctypes as C
import shlex
from dataclasses import dataclass, field

from . import libimagew as L

_FMT_BY_NAME = {"png": L.IW_FORMAT_PNG, "jpg": L.IW_FORMAT_JPEG, "jpeg": L.IW_FORMAT_JPEG, "bmp": L.IW_FORMAT_BMP,
                "tif": L.IW_FORMAT_TIFF, "tiff": L.IW_FORMAT_TIFF, "miff": L.IW_FORMAT_MIFF,
                "pnm": L.IW_FORMAT_PNM, "ppm": L.IW_FORMAT_PPM, "pgm": L.IW_FORMAT_PGM,
                "pbm": L.IW_FORMAT_PBM, "pam": L.IW_FORMAT_PAM}

# options that take no argument (reference: src/imagew-cmd.c:2183-2202)
_FLAGS = {"interlace", "bestfit", "nobestfit", "noresize", "grayscale", "condgrayscale", "nogamma", "intclamp",
          "nocslabel", "usebkgdlabel", "nobkgdlabel", "includescreen", "noincludescreen", "jpegarith",
          "bmptrns", "negate", "quiet", "nowarn", "noinfo"}


def _num_list(s: str, maxn: int) -> list[float]:
    out = []
    for tok in s.split(","):
        if len(out) >= maxn:
            break
        try:
            out.append(float(tok))
        except ValueError:
            out.append(0.0)
    return out


def _round_to_int(x: float) -> int:
    # reference: src/imagew-util.c iw_round_to_int: (x<0) ? (int)(x-0.5) : (int)(x+0.5)
    return int(x - 0.5) if x < 0 else int(x + 0.5)


def _hexdigit(d: str) -> int:
    try:
        return int(d, 16)
    except ValueError:
        return 0


def parse_color(s: str) -> tuple[list[float], bool]:
    """reference: src/imagew-cmd.c:1547-1577 parse_bkgd_color.  Returns ([r,g,b,a], ok)."""
    c = [0.0, 0.0, 0.0, 1.0]
    n = len(s)
    if n in (3, 4):
        for k in range(n):
            c[k] = _hexdigit(s[k]) / 15.0
    elif n in (6, 8):
        for k in range(n // 2):
            c[k] = (16 * _hexdigit(s[2 * k]) + _hexdigit(s[2 * k + 1])) / 255.0
    elif n in (12, 16):
        for k in range(n // 4):
            v = 0
            for d in s[4 * k:4 * k + 4]:
                v = v * 16 + _hexdigit(d)
            c[k] = v / 65535.0
    else:
        c[0], c[1], c[2] = 1.0, 0.0, 1.0
    return c, True


def _name_len(s: str) -> int:
    i = 0
    while i < len(s) and s[i].isascii() and s[i].isalpha():
        i += 1
    return i


def decode_resizetype(s: str) -> tuple[int, float, float]:
    """reference: src/imagew-cmd.c:1622-1735.  Returns (family, param1, param2)."""
    simple = {"mix": L.IW_RESIZETYPE_MIX, "nearest": L.IW_RESIZETYPE_NEAREST, "point": L.IW_RESIZETYPE_NEAREST,
              "linear": L.IW_RESIZETYPE_TRIANGLE, "triangle": L.IW_RESIZETYPE_TRIANGLE,
              "quadratic": L.IW_RESIZETYPE_QUADRATIC, "hermite": L.IW_RESIZETYPE_HERMITE,
              "box": L.IW_RESIZETYPE_BOX, "boxavg": L.IW_RESIZETYPE_BOXAVG,
              "gaussian": L.IW_RESIZETYPE_GAUSSIAN, "auto": L.IW_RESIZETYPE_AUTO, "null": L.IW_RESIZETYPE_NULL}
    if s in simple:
        return simple[s], 0.0, 0.0
    nl = _name_len(s)
    name, rest = s[:nl], s[nl:]
    if name == "lanczos":
        return L.IW_RESIZETYPE_LANCZOS, (float(rest) if rest else 3.0), 0.0
    if name in ("hann", "hanning"):
        return L.IW_RESIZETYPE_HANNING, (float(rest) if rest else 4.0), 0.0
    if name == "blackman":
        return L.IW_RESIZETYPE_BLACKMAN, (float(rest) if rest else 4.0), 0.0
    if name == "sinc":
        return L.IW_RESIZETYPE_SINC, (float(rest) if rest else 4.0), 0.0
    if s == "catrom":
        return L.IW_RESIZETYPE_CUBIC, 0.0, 0.5
    if s == "mitchell":
        return L.IW_RESIZETYPE_CUBIC, 1.0 / 3, 1.0 / 3
    if s == "bspline":
        return L.IW_RESIZETYPE_CUBIC, 1.0, 0.0
    if name == "cubic" and "," in rest:
        b, c = rest.split(",", 1)
        return L.IW_RESIZETYPE_CUBIC, float(b), float(c)
    if name == "keys":
        p2 = float(rest) if rest else 0.5
        return L.IW_RESIZETYPE_CUBIC, 1.0 - 2.0 * p2, p2
    raise ValueError(f"unknown resize type {s!r}")


_DITHERS = {  # reference: src/imagew-cmd.c:1771-1787
    "f": (2, 0), "fs": (2, 0), "o": (1, 0), "halftone": (1, 1), "r": (3, 0), "r2": (3, 1), "jjn": (2, 1),
    "stucki": (2, 2), "burkes": (2, 3), "sierra": (2, 4), "sierra3": (2, 4), "sierra2": (2, 5),
    "sierralite": (2, 6), "atkinson": (2, 7), "none": (0, 0)}

_REORIENT = {"fliph": 1, "flipv": 2, "rotate180": 3, "transpose": 4, "rotate90": 5, "rotate270": 6, "transverse": 7}


@dataclass
class Job:
    outfmt: int = L.IW_FORMAT_PNG
    dst_width_req: int = -1
    dst_height_req: int = -1
    rel_width: float | None = None
    rel_height: float | None = None
    bestfit: bool = False
    noresize: bool = False
    depth: int = 0
    depthcc: int = 0
    channel_depth: list = field(default_factory=lambda: [0] * 5)
    cs_in: tuple | None = None      # ("srgb"|"linear"|"rec709"|"gamma", gamma)
    cs_out: tuple | None = None
    alg_x: tuple | None = None      # (family, p1, p2)
    alg_y: tuple | None = None
    blur_x: tuple = (False, False, 1.0)   # (is_set, interpolate, blur)
    blur_y: tuple = (False, False, 1.0)
    dither_all: tuple | None = None
    dither_nonalpha: tuple | None = None
    dither: list = field(default_factory=lambda: [None] * 5)
    cc_all: int = 0
    cc_nonalpha: int = 0
    cc: list = field(default_factory=lambda: [0] * 5)
    apply_bkgd: bool = False
    bkgd: list | None = None
    bkgd2: list | None = None
    bkgd_checkerboard: bool = False
    check_size: int = 16
    check_origin: tuple = (0, 0)
    bkgd_label: list | None = None
    use_bkgd_label: bool = False
    no_bkgd_label: bool = False
    crop: tuple | None = None
    reorient: int = 0
    offset_h: list = field(default_factory=lambda: [0.0] * 3)
    offset_v: list = field(default_factory=lambda: [0.0] * 3)
    translate: tuple | None = None   # (x, y, src_flag)
    imagesize: tuple | None = None
    random_seed: int = 0
    randomize: bool = False
    edge_x: int = -1
    edge_y: int = -1
    gsf: int = -1
    gs_weights: list = field(default_factory=lambda: [0.0] * 3)
    grayscale: bool = False
    condgrayscale: bool = False
    no_gamma: bool = False
    intclamp: bool = False
    negate: bool = False
    bmp_trns: bool = False
    bmp_version: int = 0
    noopt: set = field(default_factory=set)
    options: list = field(default_factory=list)      # -opt name=val pairs
    disable_all_opt: bool = False                       # harness knob: iw_set_allow_opt(*, 0)
    profile_override: int | None = None                 # harness knob
    max_malloc: int | None = None                       # harness knob
    # filled by apply():
    src_width: int = 0
    src_height: int = 0
    dst_width: int = 0
    dst_height: int = 0

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def parse(cmdline, outfmt: str = "png", **knobs) -> "Job":
        args = shlex.split(cmdline) if isinstance(cmdline, str) else list(cmdline)
        j = Job(outfmt=_FMT_BY_NAME[outfmt])
        for k, v in knobs.items():
            setattr(j, k, v)
        i = 0
        while i < len(args):
            a = args[i]
            i += 1
            if not a.startswith("-"):
                continue  # a file name
            name = a.lstrip("-")
            if name in _FLAGS:
                j._flag(name)
                continue
            if i >= len(args):
                raise ValueError(f"option {a} needs an argument")
            j._opt(name, args[i])
            i += 1
        return j

    def _flag(self, n: str) -> None:
        if n == "bestfit":
            self.bestfit = True
        elif n == "nobestfit":
            self.bestfit = False
        elif n == "noresize":
            self.noresize = True
        elif n == "grayscale":
            self.grayscale = True
        elif n == "condgrayscale":
            self.condgrayscale = True
        elif n == "nogamma":
            self.no_gamma = True
        elif n == "intclamp":
            self.intclamp = True
        elif n == "usebkgdlabel":
            self.use_bkgd_label = True
        elif n == "nobkgdlabel":
            self.no_bkgd_label = True
        elif n == "negate":
            self.negate = True
        elif n == "bmptrns":
            self.bmp_trns = True

    @staticmethod
    def _w_or_h(v: str):
        if v.startswith("x"):
            return None, float(v[1:])
        return int(float(v)), None

    def _opt(self, n: str, v: str) -> None:
        if n in ("w", "width"):
            d, r = self._w_or_h(v)
            if r is None:
                self.dst_width_req = d
            else:
                self.rel_width = r
        elif n in ("h", "height"):
            d, r = self._w_or_h(v)
            if r is None:
                self.dst_height_req = d
            else:
                self.rel_height = r
        elif n in ("s", "S"):
            a, b = v.split(",", 1)
            self._opt("w", a)
            self._opt("h", b)
            self.bestfit = (n == "s")
        elif n == "opt":
            name, _, val = v.partition("=")
            self.options.append((name, val))
            if name == "bmp:version":
                self.bmp_version = 5 if val == "auto" else int(val)
        elif n == "bmpversion":
            self.options.append(("bmp:version", v))
            self.bmp_version = 5 if v == "auto" else int(v)
        elif n == "depth":
            if "," in v:
                nums = _num_list(v, 4)
                for k, x in zip((0, 1, 2, 3), nums):
                    self.channel_depth[k] = _round_to_int(x)
            else:
                self.depth = int(float(v))
        elif n == "depthcc":
            self.depthcc = int(float(v))
        elif n == "depthgray":
            self.channel_depth[L.IW_CHANNELTYPE_GRAY] = int(float(v))
        elif n == "depthalpha":
            self.channel_depth[L.IW_CHANNELTYPE_ALPHA] = int(float(v))
        elif n == "inputcs":
            self.cs_in = self._colorspace(v)
        elif n == "cs":
            self.cs_out = self._colorspace(v)
        elif n == "filter":
            self.alg_x = self.alg_y = decode_resizetype(v)
        elif n == "filterx":
            self.alg_x = decode_resizetype(v)
        elif n == "filtery":
            self.alg_y = decode_resizetype(v)
        elif n in ("blur", "blurx", "blury"):
            if v.startswith("x"):
                b = (True, True, 1.0 if len(v) == 1 else float(v[1:]))
            else:
                b = (True, False, float(v))
            if n != "blury":
                self.blur_x = b
            if n != "blurx":
                self.blur_y = b
        elif n == "dither":
            self.dither_all = _DITHERS[v]
        elif n == "dithercolor":
            self.dither_nonalpha = _DITHERS[v]
        elif n in ("ditherred", "dithergreen", "ditherblue", "ditheralpha", "dithergray"):
            k = {"ditherred": 0, "dithergreen": 1, "ditherblue": 2, "ditheralpha": 3, "dithergray": 4}[n]
            self.dither[k] = _DITHERS[v]
        elif n == "cc":
            if "," in v:
                for k, x in zip((0, 1, 2, 3), _num_list(v, 4)):
                    self.cc[k] = _round_to_int(x)
            else:
                self.cc_all = int(float(v))
        elif n == "cccolor":
            self.cc_nonalpha = int(float(v))
        elif n in ("ccred", "ccgreen", "ccblue", "ccalpha", "ccgray"):
            k = {"ccred": 0, "ccgreen": 1, "ccblue": 2, "ccalpha": 3, "ccgray": 4}[n]
            self.cc[k] = int(float(v))
        elif n == "bkgd":
            self.apply_bkgd = True
            if "," in v:
                a, b = v.split(",", 1)
                self.bkgd_checkerboard = True
                self.bkgd, _ = parse_color(a)
                self.bkgd2, _ = parse_color(b)
            else:
                self.bkgd, _ = parse_color(v)
        elif n == "bkgdlabel":
            self.bkgd_label, _ = parse_color(v)
        elif n == "checkersize":
            self.check_size = int(float(v))
        elif n == "checkerorigin":
            nums = _num_list(v, 2)
            ox, oy = self.check_origin
            if len(nums) >= 1:
                ox = _round_to_int(nums[0])
            if len(nums) >= 2:
                oy = _round_to_int(nums[1])
            self.check_origin = (ox, oy)
        elif n == "crop":
            c = [0, 0, -1, -1]
            for k, x in enumerate(_num_list(v, 4)):
                c[k] = _round_to_int(x)
            self.crop = tuple(c)
        elif n == "reorient":
            self.reorient = int(v) if v[0].isdigit() else _REORIENT[v]
        elif n in ("offsetred", "offsetgreen", "offsetblue"):
            self.offset_h[{"offsetred": 0, "offsetgreen": 1, "offsetblue": 2}[n]] = float(v)
        elif n in ("offsetvred", "offsetvgreen", "offsetvblue"):
            self.offset_v[{"offsetvred": 0, "offsetvgreen": 1, "offsetvblue": 2}[n]] = float(v)
        elif n == "offsetrb":
            self.offset_h[0] = float(v)
            self.offset_h[2] = -self.offset_h[0]
        elif n == "offsetvrb":
            self.offset_v[0] = float(v)
            self.offset_v[2] = -self.offset_v[0]
        elif n == "translate":
            src = v.startswith("s")
            nums = _num_list(v[1:] if src else v, 2)
            x = nums[0] if len(nums) >= 1 else 0.0
            y = nums[1] if len(nums) >= 2 else 0.0
            self.translate = (x, y, src)
        elif n == "imagesize":
            nums = _num_list(v, 2)
            if len(nums) == 2 and nums[0] > 0.0 and nums[1] > 0.0:
                self.imagesize = (nums[0], nums[1])
        elif n == "randseed":
            if v.startswith("r"):
                self.randomize = True
            else:
                self.random_seed = int(float(v))
        elif n == "outfmt":
            self.outfmt = _FMT_BY_NAME[v]
        elif n == "edge":
            self.edge_x = self.edge_y = self._edge(v)
        elif n == "edgex":
            self.edge_x = self._edge(v)
        elif n == "edgey":
            self.edge_y = self._edge(v)
        elif n in ("gsf", "grayscaleformula"):
            self.grayscale = True
            if v == "s":
                self.gsf = L.IW_GSF_STANDARD
            elif v == "c":
                self.gsf = L.IW_GSF_COMPATIBLE
            elif v[0] in "wv":
                nums = _num_list(v[1:], 3)
                for k, x in enumerate(nums):
                    self.gs_weights[k] = x
                self.gsf = L.IW_GSF_WEIGHTED if v[0] == "w" else L.IW_GSF_ORDERBYVALUE
            else:
                raise ValueError("unknown grayscale formula")
        elif n == "noopt":
            if v == "all":
                self.noopt |= {"grayscale", "palette", "reduceto8", "stripalpha", "binarytrns"}
            else:
                self.noopt.add(v)
        elif n in ("zipcmprlevel", "pngcmprlevel", "compress", "jpegquality", "jpegsampling", "webpquality",
                   "colortype", "page", "density", "intent", "precision", "encoding", "infmt", "sampletype"):
            pass  # codec-side; no influence on the pixel pipeline
        else:
            raise ValueError(f"unknown option -{n}")

    @staticmethod
    def _edge(v: str) -> int:
        return {"s": L.IW_EDGE_POLICY_STANDARD, "r": L.IW_EDGE_POLICY_REPLICATE,
                "t": L.IW_EDGE_POLICY_TRANSPARENT}[v[0]]

    @staticmethod
    def _colorspace(s: str):
        if s.startswith("gamma") and len(s) > 5:
            return ("gamma", float(s[5:]))
        if s == "linear":
            return ("linear", 0.0)
        if s.startswith("srgb"):
            return ("srgb", 0.0)
        if s == "rec709":
            return ("rec709", 0.0)
        raise ValueError(f"unknown colour space {s!r}")

    @staticmethod
    def _make_cs(lib: L.Lib, spec) -> L.iw_csdescr:
        cs = L.iw_csdescr()
        kind, g = spec
        if kind == "gamma":
            lib.iw_make_gamma_csdescr(C.byref(cs), g)
        elif kind == "linear":
            lib.iw_make_linear_csdescr(C.byref(cs))
        elif kind == "rec709":
            lib.iw_make_rec709_csdescr(C.byref(cs))
        else:
            lib.iw_make_srgb_csdescr_2(C.byref(cs))
        return cs

    # ------------------------------------------------------------------------------------------
    def _do_bestfit(self, asw: int, ash: int) -> None:
        """reference: src/imagew-cmd.c:427-454"""
        x = int(0.5 + ash * (self.dst_width / asw))
        if x <= self.dst_height:
            self.dst_height = x
        else:
            x = int(0.5 + asw * (self.dst_height / ash))
            if x < self.dst_width:
                self.dst_width = x
        self.dst_width = max(1, self.dst_width)
        self.dst_height = max(1, self.dst_height)

    def _figure_out_size(self) -> None:
        """reference: src/imagew-cmd.c:805-905 (square pixels assumed; density handling is codec-side)"""
        dwr, dhr = self.dst_width_req, self.dst_height_req
        relw, relh = self.rel_width, self.rel_height
        bestfit = self.bestfit
        if self.noresize:
            dwr, dhr, relw, relh, bestfit = self.src_width, self.src_height, None, None, False
        asw, ash = self.src_width, self.src_height
        self.dst_width, self.dst_height = dwr, dhr
        if self.dst_width < 0:
            self.dst_width = -1
        if self.dst_height < 0:
            self.dst_height = -1
        if self.dst_width == 0:
            self.dst_width = 1
        if self.dst_height == 0:
            self.dst_height = 1
        if relw is not None:
            self.dst_width = max(1, int(0.5 + relw * asw))
        if relh is not None:
            self.dst_height = max(1, int(0.5 + relh * ash))
        if self.dst_width == -1 and self.dst_height == -1:
            self.dst_width, self.dst_height = asw, ash
        elif self.dst_height == -1:
            self.dst_height = 1000000
            self._do_bestfit(asw, ash)
        elif self.dst_width == -1:
            self.dst_width = 1000000
            self._do_bestfit(asw, ash)
        elif bestfit:
            self._do_bestfit(asw, ash)
        self.dst_width = max(1, self.dst_width)
        self.dst_height = max(1, self.dst_height)

    def _set_bitdepth(self, lib: L.Lib, ctx) -> None:
        """reference: src/imagew-cmd.c:984-1044"""
        overall = self.depth
        if self.depthcc >= 2:
            for k in range(5):
                lib.iw_set_output_max_color_code(ctx, k, self.depthcc - 1)
            return
        cd = list(self.channel_depth)
        for k in range(5):
            overall = max(overall, cd[k])
        if not overall:
            return
        if self.outfmt == L.IW_FORMAT_MIFF:
            overall = 32 if overall <= 32 else 64
        else:
            overall = 8 if overall <= 8 else 16
        lib.iw_set_output_depth(ctx, overall)
        if self.depth:
            for k in range(5):
                if not cd[k]:
                    cd[k] = self.depth
        for k in range(5):
            if 0 < cd[k] < overall:
                lib.iw_set_output_max_color_code(ctx, k, (1 << cd[k]) - 1)

    def profile(self, lib: L.Lib) -> int:
        """reference: src/imagew-cmd.c:1213-1226"""
        if self.profile_override is not None:
            return self.profile_override
        p = lib.iw_get_profile_by_fmt(self.outfmt)
        if self.outfmt == L.IW_FORMAT_BMP:
            if self.bmp_trns:
                p |= 0x0004 | L.IW_PROFILE_TRANSPARENCY | L.IW_PROFILE_RGB8_BKGD
            if self.bmp_version >= 5:
                p |= L.IW_PROFILE_TRANSPARENCY
            p |= L.IW_PROFILE_16BPS
        return p

    def apply(self, lib: L.Lib, ctx) -> None:
        """Issue the API calls `iwcmd_run` would issue between reading the input and iw_process_image
        (reference: src/imagew-cmd.c:1108-1405).  The input image must already be set."""
        for name, val in self.options:
            lib.iw_set_option(ctx, name.encode(), val.encode())
        if self.max_malloc is not None:
            lib.iw_set_max_malloc(ctx, self.max_malloc)
        if self.random_seed != 0 or self.randomize:
            lib.iw_set_random_seed(ctx, int(self.randomize), self.random_seed)
        if self.no_gamma:
            lib.iw_set_value(ctx, L.IW_VAL_DISABLE_GAMMA, 1)
        if self.intclamp:
            lib.iw_set_value(ctx, L.IW_VAL_INT_CLAMP, 1)
        noopt = set(self.noopt)
        if self.disable_all_opt:
            noopt |= {"grayscale", "palette", "reduceto8", "stripalpha", "binarytrns"}
        for nm, code in (("grayscale", L.IW_OPT_GRAYSCALE), ("palette", L.IW_OPT_PALETTE),
                         ("reduceto8", L.IW_OPT_16_TO_8), ("stripalpha", L.IW_OPT_STRIP_ALPHA),
                         ("binarytrns", L.IW_OPT_BINARY_TRNS)):
            if nm in noopt:
                lib.iw_set_allow_opt(ctx, code, 0)
        if self.edge_x >= 0:
            lib.iw_set_value(ctx, L.IW_VAL_EDGE_POLICY_X, self.edge_x)
        if self.edge_y >= 0:
            lib.iw_set_value(ctx, L.IW_VAL_EDGE_POLICY_Y, self.edge_y)
        if self.gsf >= 0:
            lib.iw_set_value(ctx, L.IW_VAL_GRAYSCALE_FORMULA, self.gsf)
            if self.gsf in (L.IW_GSF_WEIGHTED, L.IW_GSF_ORDERBYVALUE):
                lib.iw_set_grayscale_weights(ctx, *self.gs_weights)
        if self.negate:
            lib.iw_set_value(ctx, L.IW_VAL_NEGATE_TARGET, 1)

        if self.reorient:
            lib.iw_reorient_image(ctx, self.reorient)

        lib.iw_set_output_profile(ctx, self.profile(lib))
        self._set_bitdepth(lib, ctx)

        if self.cs_in is not None:
            cs = self._make_cs(lib, self.cs_in)
            lib.iw_set_input_colorspace(ctx, C.byref(cs))
        if self.cs_out is not None:
            cs = self._make_cs(lib, self.cs_out)
            lib.iw_set_output_colorspace(ctx, C.byref(cs))

        if self.dither_all is not None:
            lib.iw_set_dither_type(ctx, L.IW_CHANNELTYPE_ALL, *self.dither_all)
        if self.dither_nonalpha is not None:
            lib.iw_set_dither_type(ctx, L.IW_CHANNELTYPE_NONALPHA, *self.dither_nonalpha)
        for k in range(5):
            if self.dither[k] is not None:
                lib.iw_set_dither_type(ctx, k, *self.dither[k])

        cc_all = 2 if self.outfmt == L.IW_FORMAT_PBM else self.cc_all
        if cc_all:
            lib.iw_set_color_count(ctx, L.IW_CHANNELTYPE_ALL, cc_all)
        if self.cc_nonalpha:
            lib.iw_set_color_count(ctx, L.IW_CHANNELTYPE_NONALPHA, self.cc_nonalpha)
        for k in range(5):
            if self.cc[k]:
                lib.iw_set_color_count(ctx, k, self.cc[k])

        grayscale, cond = self.grayscale, self.condgrayscale
        if self.outfmt in (L.IW_FORMAT_PGM, L.IW_FORMAT_PBM):
            grayscale, cond = True, False
        if cond:
            if lib.iw_get_value(ctx, L.IW_VAL_INPUT_NATIVE_GRAYSCALE):
                lib.iw_set_value(ctx, L.IW_VAL_CVT_TO_GRAYSCALE, 1)
        elif grayscale:
            lib.iw_set_value(ctx, L.IW_VAL_CVT_TO_GRAYSCALE, 1)

        for k in range(3):
            if self.offset_h[k] != 0.0:
                lib.iw_set_channel_offset(ctx, k, L.IW_DIMENSION_H, self.offset_h[k])
            if self.offset_v[k] != 0.0:
                lib.iw_set_channel_offset(ctx, k, L.IW_DIMENSION_V, self.offset_v[k])

        if self.apply_bkgd:
            b1 = list(self.bkgd)
            b2 = list(self.bkgd2) if self.bkgd2 else [0.0, 0.0, 0.0, 1.0]
            if not self.no_gamma:
                srgb = L.iw_csdescr()
                lib.iw_make_srgb_csdescr_2(C.byref(srgb))
                for k in range(3):
                    b1[k] = lib.iw_convert_sample_to_linear(b1[k], C.byref(srgb))
                    if self.bkgd_checkerboard:
                        b2[k] = lib.iw_convert_sample_to_linear(b2[k], C.byref(srgb))
            clr = L.iw_color()
            clr.c[:] = b1
            lib.iw_set_apply_bkgd_2(ctx, C.byref(clr))
            if self.bkgd_checkerboard:
                clr2 = L.iw_color()
                clr2.c[:] = b2
                lib.iw_set_bkgd_checkerboard_2(ctx, self.check_size, C.byref(clr2))
                lib.iw_set_bkgd_checkerboard_origin(ctx, *self.check_origin)

        if self.use_bkgd_label:
            lib.iw_set_value(ctx, L.IW_VAL_USE_BKGD_LABEL, 1)
        if self.no_bkgd_label:
            lib.iw_set_value(ctx, L.IW_VAL_NO_BKGD_LABEL, 1)
        elif self.bkgd_label is not None:
            srgb = L.iw_csdescr()
            lib.iw_make_srgb_csdescr_2(C.byref(srgb))
            lab = list(self.bkgd_label)
            for k in range(3):
                lab[k] = lib.iw_convert_sample_to_linear(lab[k], C.byref(srgb))
            clr = L.iw_color()
            clr.c[:] = lab
            lib.iw_set_output_bkgd_label_2(ctx, C.byref(clr))

        self.src_width = lib.iw_get_value(ctx, L.IW_VAL_INPUT_WIDTH)
        self.src_height = lib.iw_get_value(ctx, L.IW_VAL_INPUT_HEIGHT)
        crop = None
        if self.crop is not None:
            cx, cy, cw, ch = self.crop
            cx, cy = max(cx, 0), max(cy, 0)
            cx = min(cx, self.src_width - 1)
            cy = min(cy, self.src_height - 1)
            if cw < 0 or cw > self.src_width - cx:
                cw = self.src_width - cx
            if ch < 0 or ch > self.src_height - cy:
                ch = self.src_height - cy
            cw, ch = max(cw, 1), max(ch, 1)
            crop = (cx, cy, cw, ch)
            self.src_width, self.src_height = cw, ch

        self._figure_out_size()

        if self.translate is not None:
            tx, ty, src = self.translate
            if src:
                if tx != 0.0:
                    tx *= self.dst_width / self.src_width
                if ty != 0.0:
                    ty *= self.dst_height / self.src_height
            lib.iw_set_value_dbl(ctx, L.IW_VAL_TRANSLATE_X, tx)
            lib.iw_set_value_dbl(ctx, L.IW_VAL_TRANSLATE_Y, ty)

        alg_x, alg_y = self.alg_x, self.alg_y
        if self.blur_x[0] and alg_x is None:
            alg_x = (L.IW_RESIZETYPE_GAUSSIAN, 0.0, 0.0)
        if self.blur_y[0] and alg_y is None:
            alg_y = (L.IW_RESIZETYPE_GAUSSIAN, 0.0, 0.0)
        if alg_x is not None:
            blur = self.blur_x[2]
            if self.blur_x[1] and self.dst_width < self.src_width:
                blur *= self.dst_width / self.src_width
            lib.iw_set_resize_alg(ctx, L.IW_DIMENSION_H, alg_x[0], blur, alg_x[1], alg_x[2])
        if alg_y is not None:
            blur = self.blur_y[2]
            if self.blur_y[1] and self.dst_height < self.src_height:
                blur *= self.dst_height / self.src_height
            lib.iw_set_resize_alg(ctx, L.IW_DIMENSION_V, alg_y[0], blur, alg_y[1], alg_y[2])

        lib.iw_set_output_canvas_size(ctx, self.dst_width, self.dst_height)
        if self.imagesize is not None:
            lib.iw_set_output_image_size(ctx, *self.imagesize)
        if crop is not None:
            lib.iw_set_input_crop(ctx, *crop)
