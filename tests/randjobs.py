"""Seeded random (input image, command line) generator for differential tests against the reference."""
from __future__ import annotations

import numpy as np

from imageworsener_b200 import libimagew as L

FILTERS = ["catrom", "mitchell", "bspline", "lanczos", "lanczos2", "lanczos5", "hanning", "blackman", "sinc3",
           "gaussian", "triangle", "quadratic", "hermite", "box", "boxavg", "mix", "nearest", "cubic0.3,0.2",
           "keys0.4", "auto"]
DITHERS = ["none", "o", "halftone", "r", "r2", "f", "jjn", "stucki", "burkes", "sierra", "sierra2", "sierralite",
           "atkinson"]


def make(seed: int, max_in: int = 60, max_out: int = 90, allow_errdiff: bool = True):
    r = np.random.default_rng(seed)
    imgtype = [L.IW_IMGTYPE_GRAY, L.IW_IMGTYPE_GRAYA, L.IW_IMGTYPE_RGB, L.IW_IMGTYPE_RGBA][r.integers(0, 4)]
    nch = L.imgtype_num_channels(imgtype)
    depth = [8, 8, 16][r.integers(0, 3)]
    if imgtype == L.IW_IMGTYPE_GRAY and r.random() < 0.25:
        depth = [1, 2, 4][r.integers(0, 3)]
    w, h = int(r.integers(1, max_in + 1)), int(r.integers(1, max_in + 1))
    bpr = (w * nch * depth + 7) // 8 + int(r.integers(0, 3))       # sometimes padded rows
    kind = r.integers(0, 3)
    if kind == 0:
        rows = r.integers(0, 256, size=(h, bpr), dtype=np.uint8)      # noise
    elif kind == 1:                                                   # smooth-ish gradient bytes
        yy, xx = np.mgrid[0:h, 0:bpr]
        rows = ((xx * 7 + yy * 5 + r.integers(0, 50)) % 256).astype(np.uint8)
    else:                                                             # flat regions (exercises exact codes)
        rows = np.full((h, bpr), int(r.integers(0, 256)), np.uint8)
        rows[: h // 2, : bpr // 2] = int(r.integers(0, 256))
    if nch in (2, 4) and r.random() < 0.4:                           # make some pixels fully transparent
        flat = rows.reshape(h, bpr)
        bps = depth // 8 if depth >= 8 else 1
        for y in range(h):
            for x in range(w):
                if r.random() < 0.3:
                    z = (x * nch + nch - 1) * bps
                    flat[y, z:z + bps] = 0
    inp = {"imgtype": imgtype, "bit_depth": depth, "width": w, "height": h, "rows": rows}
    if r.random() < 0.3:
        inp["cs"] = [("linear", 0.0), ("gamma", 2.2), ("rec709", 0.0), ("gamma", 1.8)][r.integers(0, 4)]

    ow, oh = int(r.integers(1, max_out + 1)), int(r.integers(1, max_out + 1))
    if r.random() < 0.15:
        ow, oh = w, h
    a = [f"-width {ow} -height {oh}"]
    if r.random() < 0.8:
        if r.random() < 0.7:
            a.append(f"-filter {FILTERS[r.integers(0, len(FILTERS))]}")
        else:
            a.append(f"-filterx {FILTERS[r.integers(0, len(FILTERS))]} -filtery {FILTERS[r.integers(0, len(FILTERS))]}")
    if r.random() < 0.15:
        a.append(f"-blur {r.choice([0.7, 1.3, 2.0])}")
    if r.random() < 0.4:
        a.append(f"-edge {r.choice(['s', 'r', 't'])}")
    elif r.random() < 0.1:
        a.append(f"-edgex {r.choice(['s', 'r', 't'])} -edgey {r.choice(['s', 'r', 't'])}")
    d = DITHERS[r.integers(0, len(DITHERS))] if r.random() < 0.6 else "none"
    if not allow_errdiff and d not in ("none", "o", "halftone", "r", "r2"):
        d = "o"
    if d != "none":
        a.append(f"-dither {d}")
        if r.random() < 0.5:
            a.append(f"-randseed {int(r.integers(-5, 1000))}")
    if r.random() < 0.3:
        a.append(f"-cc {int(r.choice([2, 3, 4, 6, 17, 100]))}")
    elif r.random() < 0.1:
        a.append("-cc 3,10,5,4")
    if r.random() < 0.3:
        a.append("-depth 16")
    elif r.random() < 0.1:
        a.append(f"-depth {r.choice(['5', '3,4,6', '6,8,5,7'])}")
    if r.random() < 0.2:
        a.append(f"-cs {r.choice(['linear', 'gamma1.5', 'rec709', 'srgb'])}")
    if r.random() < 0.1:
        a.append("-nogamma")
    if r.random() < 0.1:
        a.append("-intclamp")
    if r.random() < 0.25:
        a.append("-grayscale")
        if r.random() < 0.5:
            a.append(f"-gsf {r.choice(['c', 's', 'w3,7,1', 'v.4,.1,.4'])}")
    if r.random() < 0.3:
        if r.random() < 0.5:
            a.append(f"-bkgd {r.choice(['e42', '808080', 'e42d', '00ff5550'])}")
        else:
            a.append(f"-bkgd {r.choice(['e42,0f5', '876,554433', 'e42d,00ff5550'])} -checkersize {int(r.integers(1, 9))}"
                     f" -checkerorigin {int(r.integers(0, 5))},{int(r.integers(0, 5))}")
    if r.random() < 0.15:
        a.append(f"-translate {r.choice(['3,4', '-4,-3', '0.5,0.25', 's2,1'])}")
    if r.random() < 0.1:
        a.append(f"-imagesize {ow * 0.8:.2f},{oh * 0.7:.2f}")
    if r.random() < 0.1:
        a.append(f"-offsetred {r.choice(['.333', '-1.5'])} -offsetvblue {r.choice(['.6', '-0.2'])}")
    if r.random() < 0.15:
        a.append(f"-reorient {int(r.integers(1, 8))}")
    if r.random() < 0.15 and w > 4 and h > 4:
        a.append(f"-crop {int(r.integers(0, w // 2))},{int(r.integers(0, h // 2))},{int(r.integers(1, w))},{int(r.integers(1, h))}")
    if r.random() < 0.1:
        a.append("-negate")
    outfmt = "png"
    q = r.random()
    if q < 0.1:
        outfmt = "ppm"
    elif q < 0.15:
        outfmt = "pam"
    elif q < 0.2:
        outfmt = "bmp"
    elif q < 0.25:
        outfmt = "miff"
    return inp, " ".join(a), outfmt


def make_float(seed: int, **kw):
    """A random job whose input image holds 32-bit floating-point samples (src/imagew.h:387-390: native byte order),
    in a linear or a non-linear colourspace.  Sample values lie in [0, 1.25]: mostly code/255 of the integer job's
    bytes, some beyond 1.0 (HDR-ish input; the pipeline clamps late, src/imagew-main.c:684-685)."""
    inp, args, outfmt = make(seed, allow_errdiff=kw.pop("allow_errdiff", False), **kw)
    r = np.random.default_rng(seed + 77_000)
    nch = L.imgtype_num_channels(inp["imgtype"])
    w, h = inp["width"], inp["height"]
    src = r.integers(0, 256, size=(h, w * nch)).astype(np.float32) / np.float32(255.0)
    if r.random() < 0.5:
        src *= np.float32(1.25)
    if r.random() < 0.3:
        src[: max(1, h // 3)] = np.float32(r.choice([0.0, 1.0, 0.5]))
    if nch in (2, 4):                                                # alpha stays in [0, 1]
        src.reshape(h, w, nch)[:, :, nch - 1] = np.minimum(src.reshape(h, w, nch)[:, :, nch - 1], np.float32(1.0))
    pad = int(r.integers(0, 2)) * 4
    rows = np.zeros((h, w * nch * 4 + pad), np.uint8)
    rows[:, : w * nch * 4] = src.view(np.uint8).reshape(h, w * nch * 4)
    inp = dict(inp, bit_depth=32, sampletype=L.IW_SAMPLETYPE_FLOATINGPOINT, rows=rows)
    inp.pop("maxcodes", None)
    inp["cs"] = [None, ("linear", 0.0), ("gamma", 2.2), ("rec709", 0.0), ("srgb", 0.0)][r.integers(0, 5)]
    if inp["cs"] is None:
        inp.pop("cs")                                                # the context default: sRGB
    return inp, args, outfmt
