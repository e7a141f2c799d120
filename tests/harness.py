"""Shared helpers for the parity tests: drive one libimagew implementation with a job and canonicalise
its output so that two implementations (or an implementation and a golden vector) can be compared."""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from imageworsener_b200 import libimagew as L
from imageworsener_b200.jobs import Job

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")


def make_cs(lib: L.Lib, spec):
    return Job._make_cs(lib, spec)


def run_job(lib: L.Lib, inp: dict, job: Job, backend_opts=None, keep_ctx=False):
    """inp: dict(imgtype, bit_depth, width, height, rows[H,bpr] uint8, optional cs, bkgd_label,
    native_grayscale, sampletype, maxcodes{channel:code}).  Returns dict(ok, err, warnings, img fields, pixels)."""
    ctx = lib.iw_create_context(None)
    assert ctx
    warnings = []

    @L.WARNFN
    def _warn(_ctx, msg):
        warnings.append(msg.decode("utf-8", "replace"))

    lib.iw_set_warning_fn(ctx, _warn)
    try:
        lib.set_input_from_numpy(ctx, inp["rows"], inp["imgtype"], inp["bit_depth"], inp["width"], inp["height"],
                                 sampletype=inp.get("sampletype", 0),
                                 native_grayscale=inp.get("native_grayscale", 0))
        if inp.get("cs") is not None:
            cs = make_cs(lib, tuple(inp["cs"]))
            lib.iw_set_input_colorspace(ctx, C.byref(cs))
        if inp.get("bkgd_label") is not None:
            lib.iw_set_input_bkgd_label(ctx, *inp["bkgd_label"])
        for ch, code in (inp.get("maxcodes") or {}).items():
            lib.iw_set_input_max_color_code(ctx, int(ch), int(code))
        if lib.is_product and backend_opts:
            for k, v in backend_opts.items():
                lib.iw_set_option(ctx, k.encode(), str(v).encode())
        job.apply(lib, ctx)
        ok = lib.iw_process_image(ctx)
        res = {"ok": bool(ok), "warnings": warnings, "err": None}
        if not ok:
            res["err"] = lib.errmsg(ctx)
            return res
        img, px = lib.get_output_numpy(ctx)
        cs = L.iw_csdescr()
        lib.iw_get_output_colorspace(ctx, C.byref(cs))
        res.update(imgtype=img.imgtype, bit_depth=img.bit_depth, sampletype=img.sampletype, width=img.width,
                   height=img.height, bpr=img.bpr, pixels=px, reduced_maxcolors=img.reduced_maxcolors,
                   maxcolorcode=list(img.maxcolorcode), has_bkgdlabel=img.has_bkgdlabel,
                   bkgdlabel=list(img.bkgdlabel.c), out_cstype=cs.cstype, out_gamma=cs.gamma)
        return res
    finally:
        if not keep_ctx:
            lib.iw_destroy_context(ctx)


def samples_of(res: dict) -> np.ndarray:
    """(H, W, nch) integer (or float32) samples of an output result."""
    h, w = res["height"], res["width"]
    nch = L.imgtype_num_channels(res["imgtype"])
    px = res["pixels"]
    if res["sampletype"] == L.IW_SAMPLETYPE_FLOATINGPOINT:
        return px[:, :w * nch * 4].copy().view(np.float32).reshape(h, w, nch)
    if res["bit_depth"] == 8:
        return px[:, :w * nch].reshape(h, w, nch).astype(np.uint32)
    if res["bit_depth"] == 16:
        r = px[:, :w * nch * 2].reshape(h, w, nch, 2).astype(np.uint32)
        return r[..., 0] * 256 + r[..., 1]
    raise ValueError(f"unexpected output depth {res['bit_depth']}")


def canonical_rgba(res: dict) -> np.ndarray:
    """Output as (H, W, 4) uint32 RGBA in raw code units (missing alpha = max code of the container)."""
    s = samples_of(res)
    h, w, nch = s.shape
    mx = (1 << res["bit_depth"]) - 1
    out = np.zeros((h, w, 4), np.uint32)
    out[:, :, 3] = mx
    t = res["imgtype"]
    if t == L.IW_IMGTYPE_GRAY:
        out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = s[:, :, 0]
    elif t == L.IW_IMGTYPE_GRAYA:
        out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = s[:, :, 0]
        out[:, :, 3] = s[:, :, 1]
    elif t == L.IW_IMGTYPE_RGB:
        out[:, :, :3] = s
    else:
        out[:] = s
    return out


def load_golden():
    """Returns list of case dicts (see tests/golden/make_golden.py)."""
    meta = json.load(open(os.path.join(GOLDEN_DIR, "golden_cases.json")))
    blob = np.load(os.path.join(GOLDEN_DIR, "golden_cases.npz"))
    cases = []
    for m in meta["cases"]:
        c = dict(m)
        c["input"] = dict(m["input"])
        c["input"]["rows"] = blob[m["name"] + "__in"]
        c["expected"] = blob[m["name"] + "__exp"].astype(np.uint32)
        cases.append(c)
    return cases


def _our_channel_max(res: dict) -> np.ndarray:
    """Max code of each canonical RGBA channel of an output (reduced max codes if the output uses them)."""
    mx = (1 << res["bit_depth"]) - 1
    m = np.array([mx, mx, mx, mx], np.uint64)
    if res["reduced_maxcolors"]:
        mc = res["maxcolorcode"]
        if res["imgtype"] in (L.IW_IMGTYPE_GRAY, L.IW_IMGTYPE_GRAYA):
            m[:3] = mc[L.IW_CHANNELTYPE_GRAY]
        else:
            m[:3] = mc[:3]
        if res["imgtype"] in (L.IW_IMGTYPE_GRAYA, L.IW_IMGTYPE_RGBA):
            m[3] = mc[L.IW_CHANNELTYPE_ALPHA]
    return m


def compare_with_golden(res: dict, case: dict):
    """Number of differing samples between an output and a golden vector, both as canonical RGBA.  The
    golden files went through the reference's lossless post-pass optimiser (16->8 bit, gray, palette, 1-bit)
    and, for PNM/PAM, carry their own MAXVAL, so samples are compared as exact fractions code/max:
    ours*exp_max == exp*our_max."""
    ours = canonical_rgba(res).astype(np.uint64)
    exp = case["expected"].astype(np.uint64)
    if ours.shape != exp.shape:
        return -1
    omax = _our_channel_max(res)
    emax = np.full(4, case["exp_maxval"], np.uint64)
    if res["imgtype"] in (L.IW_IMGTYPE_GRAY, L.IW_IMGTYPE_RGB):
        ours[:, :, 3] = omax[3]          # no alpha in our output: opaque
    if case["exp_alpha_missing"]:
        exp[:, :, 3] = emax[3]
    diff = ours * emax != exp * omax
    # colour of fully transparent pixels is not comparable: the reference's (out-of-scope) binary-transparency
    # optimiser recolours them to its colour key; both sides agreeing on alpha==0 is what matters
    both_clear = (ours[:, :, 3] == 0) & (exp[:, :, 3] == 0)
    diff[both_clear, :3] = False
    return int(np.count_nonzero(diff))


def xorshift_bytes(n: int, seed: int = 0x9E3779B97F4A7C15) -> np.ndarray:
    """The survey's synthetic-noise generator: xorshift64, byte = state>>24, vectorised by
    running 64 independent lanes is NOT equivalent, so this is the plain sequential definition -- use only
    for small n.  For large buffers use imageworsener_b200.synth."""
    out = np.empty(n, np.uint8)
    s = seed
    m = (1 << 64) - 1
    for i in range(n):
        s ^= (s << 13) & m
        s ^= s >> 7
        s ^= (s << 17) & m
        out[i] = (s >> 24) & 255
    return out
