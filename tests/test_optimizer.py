"""The post-pass optimiser (csrc/optimize.h, optimize.cpp; reference: iwpvt_optimize_image, src/imagew-opt.c:1166-1314)
with the reference's DEFAULT settings (every iw_set_allow_opt on): the image handed out by iw_get_output_image -- type,
depth, row stride, every byte, palette, colour key, background label -- must equal the reference's.

Noise never triggers a reduction, so the inputs here are built to: few colours (palette 1/2/4/8 bits), gray RGB, opaque
or binary alpha, 16-bit images whose samples are 257 * k (reducible to 8 bits), colour-key candidates around 192.
"""
import ctypes as C

import numpy as np
import pytest

import harness
import randjobs
from imageworsener_b200 import libimagew as L
from imageworsener_b200.jobs import Job

RGB, GRAY, RGBA, GA = L.IW_IMGTYPE_RGB, L.IW_IMGTYPE_GRAY, L.IW_IMGTYPE_RGBA, L.IW_IMGTYPE_GRAYA


class Palette(C.Structure):
    _fields_ = [("num_entries", C.c_int), ("entry", C.c_uint8 * 1024)]


def run_with_palette(lib, inp, job):
    lib.dll.iw_get_output_palette.restype = C.POINTER(Palette)
    lib.dll.iw_get_output_palette.argtypes = [C.c_void_p]
    ctx = lib.iw_create_context(None)
    try:
        lib.set_input_from_numpy(ctx, inp["rows"], inp["imgtype"], inp["bit_depth"], inp["width"], inp["height"])
        if inp.get("bkgd_label") is not None:
            lib.iw_set_input_bkgd_label(ctx, *inp["bkgd_label"])
        job.apply(lib, ctx)
        if not lib.iw_process_image(ctx):
            return {"ok": False, "err": lib.errmsg(ctx)}
        img, px = lib.get_output_numpy(ctx)
        pal = lib.dll.iw_get_output_palette(ctx)
        palette = None
        if pal:
            n = pal.contents.num_entries
            palette = bytes(pal.contents.entry[: 4 * n])
        return {"ok": True, "imgtype": img.imgtype, "bit_depth": img.bit_depth, "width": img.width, "height": img.height, "bpr": img.bpr,
                "pixels": px, "palette": palette, "has_colorkey_trns": img.has_colorkey_trns, "colorkey": list(img.colorkey),
                "has_bkgdlabel": img.has_bkgdlabel, "bkgdlabel": list(img.bkgdlabel.c)}
    finally:
        lib.iw_destroy_context(ctx)


def assert_same(a, b, tag):
    assert a["ok"] == b["ok"], (tag, a.get("err"), b.get("err"))
    if not a["ok"]:
        return
    for k in ("imgtype", "bit_depth", "width", "height", "bpr", "has_colorkey_trns", "has_bkgdlabel"):
        assert a[k] == b[k], (tag, k, a[k], b[k])
    if a["has_colorkey_trns"]:
        assert a["colorkey"] == b["colorkey"], tag
    if a["has_bkgdlabel"]:
        assert a["bkgdlabel"] == b["bkgdlabel"], tag
    assert a["palette"] == b["palette"], (tag, "palette")
    nd = int(np.count_nonzero(a["pixels"] != b["pixels"]))
    assert nd == 0, (tag, f"{nd} differing bytes")


def crafted(seed):
    """(input dict, command line, output format) built to trigger the optimiser."""
    r = np.random.default_rng(seed)
    kind = seed % 16
    w, h = int(r.integers(8, 70)), int(r.integers(8, 60))
    depth = 16 if seed % 3 == 0 else 8
    bps = depth // 8

    def expand(px):                               # px: (h, w, nch) 8-bit values -> rows at `depth` (16-bit = 257 * v: reducible)
        if bps == 1:
            return px.reshape(h, -1).astype(np.uint8)
        return np.repeat(px.reshape(h, -1), 2, axis=1).astype(np.uint8)

    levels = r.choice(256, size=int(r.integers(2, 7)), replace=False)
    if kind in (0, 1):                            # few colours, RGB
        typ, nch = RGB, 3
        idx = r.integers(0, len(levels), size=(h, w, 3))
        px = levels[idx]
    elif kind == 2:                               # gray RGB
        typ, nch = RGB, 3
        g = levels[r.integers(0, len(levels), size=(h, w, 1))]
        px = np.repeat(g, 3, axis=2)
    elif kind == 3:                               # two gray shades 0 / 255 (gray1)
        typ, nch = GRAY, 1
        px = (r.integers(0, 2, size=(h, w, 1)) * 255)
    elif kind == 4:                               # gray shades that are multiples of 17 (gray4) or 85 (gray2)
        typ, nch = GRAY, 1
        f = 85 if seed % 2 else 17
        px = r.integers(0, 255 // f + 1, size=(h, w, 1)) * f
    elif kind == 5:                               # RGBA, opaque everywhere
        typ, nch = RGBA, 4
        px = r.integers(0, 256, size=(h, w, 4))
        px[:, :, 3] = 255
    elif kind == 6:                               # RGBA, binary alpha, colours near the 192 key
        typ, nch = RGBA, 4
        px = r.integers(0, 256, size=(h, w, 4))
        px[:, :, 1:3] = 192
        px[:, :, 0] = r.choice([190, 191, 192, 193, 50], size=(h, w))
        px[:, :, 3] = r.integers(0, 2, size=(h, w)) * 255
    elif kind == 7:                               # gray + alpha, binary alpha
        typ, nch = GA, 2
        px = r.integers(0, 256, size=(h, w, 2))
        px[:, :, 0] = r.choice([0, 64, 192, 255], size=(h, w))
        px[:, :, 1] = r.integers(0, 2, size=(h, w)) * 255
    elif kind == 8:                               # gray RGBA with partial alpha
        typ, nch = RGBA, 4
        g = r.integers(0, 256, size=(h, w, 1))
        px = np.concatenate([g, g, g, r.integers(0, 256, size=(h, w, 1))], axis=2)
    elif kind == 9:                               # few colours with alpha (palette with transparency)
        typ, nch = RGBA, 4
        cols = r.integers(0, 256, size=(5, 4))
        cols[0, 3] = 0
        cols[1, 3] = 255
        px = cols[r.integers(0, 5, size=(h, w))]
    elif kind == 10:                              # gray + alpha, opaque
        typ, nch = GA, 2
        px = r.integers(0, 256, size=(h, w, 2))
        px[:, :, 1] = 255
    elif kind == 11:                              # 3 gray shades incl. transparency in RGBA (gray palette + key shade)
        typ, nch = RGBA, 4
        g = r.choice([0, 255], size=(h, w, 1))
        a = r.integers(0, 2, size=(h, w, 1)) * 255
        px = np.concatenate([g, g, g, a], axis=2)
    elif kind in (12, 13):                        # many colours, binary alpha: colour-key transparency (RGB + key)
        typ, nch = RGBA, 4
        w, h = 40, 30
        px = r.integers(0, 256, size=(h, w, 4))
        px[::2, :, 1:3] = 192                     # some pixels sit on the key colours' green / blue
        px[:, :, 3] = r.integers(0, 2, size=(h, w)) * 255
    else:                                         # many gray shades + binary alpha: gray + key
        typ, nch = GA, 2
        w, h = 40, 30
        px = r.integers(0, 256, size=(h, w, 2))
        if kind == 15:
            px[:, :, 0] = r.choice(np.delete(np.arange(256), [192, 191, 77]), size=(h, w))
        px[:, :, 1] = r.integers(0, 2, size=(h, w)) * 255
    rows = expand(px)
    if depth == 16 and kind >= 12 and seed % 2:   # real 16-bit precision: the 16-bit key searches
        rows = rows.copy()
        rows[:, 1::2] = r.integers(0, 256, size=rows[:, 1::2].shape)
        a0 = (nch - 1) * 2
        for y in range(h):
            for x in range(w):
                z = x * nch * 2 + a0
                rows[y, z + 1] = rows[y, z]        # alpha stays 0 or 65535
        rows[::3, 0::2] = 192                      # candidates for the 16-bit keys (high bytes 192)
    inp = {"imgtype": typ, "bit_depth": depth, "width": w, "height": h, "rows": rows}
    if seed % 5 == 0:
        inp["bkgd_label"] = [(0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (0.2, 0.4, 0.6)][(seed // 5) % 3]
    # keep the colours: same size + nearest (or pure copies), so the optimiser sees what was built
    args = f"-width {w} -height {h} -filter nearest -nogamma"
    if depth == 16:
        args += " -depth 16"
    if seed % 7 == 3:
        args = f"-width {w * 2} -height {h * 2} -filter nearest -nogamma" + (" -depth 16" if depth == 16 else "")
    outfmt = ["png", "png", "png", "bmp", "pam"][seed % 5]
    return inp, args, outfmt


@pytest.mark.parametrize("seed", range(240))
def test_host_logic_optimiser_on_crafted_images(emu, ref, seed):
    inp, args, outfmt = crafted(seed)
    a = run_with_palette(emu, inp, Job.parse(args, outfmt=outfmt))
    b = run_with_palette(ref, inp, Job.parse(args, outfmt=outfmt))
    assert_same(a, b, (seed, args, outfmt))


@pytest.mark.parametrize("seed", range(9000, 9150))
def test_host_logic_optimiser_on_random_jobs(emu, ref, seed):
    inp, args, outfmt = randjobs.make(seed)
    a = run_with_palette(emu, inp, Job.parse(args, outfmt=outfmt))
    b = run_with_palette(ref, inp, Job.parse(args, outfmt=outfmt))
    assert_same(a, b, (seed, args, outfmt))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(240))
def test_cuda_optimiser_on_crafted_images(prod, ref, seed):
    inp, args, outfmt = crafted(seed)
    a = run_with_palette(prod, inp, Job.parse(args, outfmt=outfmt))
    b = run_with_palette(ref, inp, Job.parse(args, outfmt=outfmt))
    assert_same(a, b, (seed, args, outfmt))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(9000, 9200))
def test_cuda_optimiser_on_random_jobs(prod, ref, seed):
    inp, args, outfmt = randjobs.make(seed)
    a = run_with_palette(prod, inp, Job.parse(args, outfmt=outfmt))
    b = run_with_palette(ref, inp, Job.parse(args, outfmt=outfmt))
    assert_same(a, b, (seed, args, outfmt))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 3, 6, 9, 11, 14, 18, 21])
def test_cuda_optimiser_with_bands_on_several_devices(prod, ref, seed):
    """cuda:devices: every band is scanned on its device, the host merges the scans, all bands are rewritten alike."""
    inp, args, outfmt = crafted(seed)
    inp = dict(inp)
    reps = 200 // inp["height"] + 2               # tall enough for several bands of 32 rows
    inp["rows"] = np.tile(inp["rows"], (reps, 1))
    inp["height"] *= reps
    args = args.replace(f"-height {inp['height'] // reps}", f"-height {inp['height']}")
    a = run_with_palette_opts(prod, inp, Job.parse(args, outfmt=outfmt), {"cuda:devices": "0,0,0"})
    b = run_with_palette(ref, inp, Job.parse(args, outfmt=outfmt))
    assert_same(a, b, (seed, args, outfmt))


def run_with_palette_opts(lib, inp, job, opts):
    orig = job.apply

    def apply(l, ctx):
        for k, v in opts.items():
            l.iw_set_option(ctx, k.encode(), str(v).encode())
        orig(l, ctx)
    job.apply = apply
    return run_with_palette(lib, inp, job)
