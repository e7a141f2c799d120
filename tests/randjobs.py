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
