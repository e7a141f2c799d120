"""Soak test of the fast path against the reference build (on the GPU box): many seeded near-integer geometries
plus full-size 8K images with every fast dither.  Prints one summary line per group; exit code 1 on any difference.

    python tools/soak.py [n_random_cases]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from imageworsener_b200 import configs, libimagew as L  # noqa: E402
import oracle  # noqa: E402  (test infrastructure: the reference build / port)
from test_stream2 import _enlargement_case, _near_integer_case  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    prod, ref = L.load_product(), oracle.load_reference()
    bad = samples = 0
    t0 = time.time()
    for seed in range(5000, 5000 + n):
        cfg = _near_integer_case(seed)
        nch = L.imgtype_num_channels(cfg.imgtype)
        tight = configs.synth_rows(prod, cfg, seed=0x777 + seed)
        bpr = (cfg.in_w * nch + 15) // 16 * 16
        rows = np.full((cfg.in_h, bpr), 0x5A, np.uint8)
        rows[:, : cfg.in_w * nch] = tight
        _, b = configs.run_host(ref, cfg, tight)
        for opts in ({}, {"cuda:uniform": 0}, {"cuda:gen": 1}):
            _, a = configs.run_host(prod, cfg, rows, options=opts)
            nd = int(np.count_nonzero(a != b))
            samples += a.size
            if nd:
                bad += 1
                print(f"DIFF seed {seed} {opts} {cfg.args!r} {cfg.in_w}x{cfg.in_h}: {nd} bytes", flush=True)
    print(f"random near-integer geometries: {n} cases x 3 variants, {samples} output bytes compared, {bad} cases differ, {time.time() - t0:.0f} s", flush=True)
    badu = samplesu = 0
    for seed in range(7000, 7000 + n):
        cfg = _enlargement_case(seed)
        nch = L.imgtype_num_channels(cfg.imgtype)
        tight = configs.synth_rows(prod, cfg, seed=0x555 + seed)
        bpr = (cfg.in_w * nch + 15) // 16 * 16
        rows = np.full((cfg.in_h, bpr), 0x5A, np.uint8)
        rows[:, : cfg.in_w * nch] = tight
        _, b = configs.run_host(ref, cfg, tight)
        for opts in ({}, {"cuda:gen": 1}):
            _, a = configs.run_host(prod, cfg, rows, options=opts)
            nd = int(np.count_nonzero(a != b))
            samplesu += a.size
            if nd:
                badu += 1
                print(f"DIFFUP seed {seed} {opts} {cfg.args!r} {cfg.in_w}x{cfg.in_h}: {nd} bytes", flush=True)
    print(f"random enlargements (12 / 16 open windows, gray as RGB): {n} cases x 2 variants, {samplesu} output bytes compared, {badu} cases differ", flush=True)
    bad += badu
    # the same geometries with 16-bit big-endian input (rows of 16-bit samples padded to 16 bytes)
    bad16 = samples16 = 0
    for seed in range(9000, 9000 + n // 2):
        c8 = _near_integer_case(seed)
        cfg = configs.Config(c8.name, c8.in_w, c8.in_h, c8.imgtype, 16, c8.out_w, c8.out_h, c8.args,
                             input_cs=("gamma", 2.2) if seed % 3 == 0 else None)
        nch = L.imgtype_num_channels(cfg.imgtype)
        tight = configs.synth_rows(prod, cfg, seed=0x999 + seed)
        bpr = (cfg.in_w * nch * 2 + 15) // 16 * 16
        rows = np.full((cfg.in_h, bpr), 0x5A, np.uint8)
        rows[:, : cfg.in_w * nch * 2] = tight
        _, b = configs.run_host(ref, cfg, tight)
        for opts in ({}, {"cuda:uniform": 0}):
            _, a = configs.run_host(prod, cfg, rows, options=opts)
            nd = int(np.count_nonzero(a != b))
            samples16 += a.size
            if nd:
                bad16 += 1
                print(f"DIFF16 seed {seed} {opts} {cfg.args!r} {cfg.in_w}x{cfg.in_h}: {nd} bytes", flush=True)
    print(f"16-bit input: {n // 2} cases x 2 variants, {samples16} output bytes compared, {bad16} cases differ", flush=True)
    bad += bad16
    # images with alpha on the gather-form kernels, most of them with 16-bit output through the level model
    # (csrc/estimate.h).  Inputs are image-like rather than noise -- smooth gradients with hard edges, exact 0 / max
    # samples and alpha -- so that undershoot, saturation, zero alpha and whole warps of equal values all occur.
    bada = samplesa = 0
    filters = ["mitchell", "catrom", "lanczos2", "lanczos", "hermite", "triangle", "gaussian", "bspline"]
    for seed in range(11000, 11000 + n):
        r = np.random.default_rng(seed)
        nch = 4 if seed % 3 else 2
        depth = 16 if seed % 4 else 8
        iw, ih = int(r.integers(40, 400)), int(r.integers(40, 300))
        f = float(r.choice([0.13, 0.25, 0.4, 0.55, 0.8, 1.0, 1.5, 2.0, 3.0, 4.0])) * float(r.uniform(0.97, 1.03))
        ow, oh = max(8, int(iw * f)), max(8, int(ih * f * float(r.uniform(0.9, 1.1))))
        out16 = seed % 5 != 0
        bk = "%02x%02x%02x" % tuple(int(v) for v in r.integers(0, 256, 3))
        cs = ["", "", "-cs rec709", "-cs srgb", "-cs gamma2.0"][seed % 5]
        dith = "" if seed % 7 else " -dither o"
        args = f"-width {ow} -height {oh} -filter {filters[seed % len(filters)]} -bkgd {bk} {cs}{dith}" + (" -depth 16" if out16 else "")
        if nch == 2:
            args += " -grayscale"
        cfg = configs.Config(f"alpha{seed}", iw, ih, L.IW_IMGTYPE_RGBA if nch == 4 else L.IW_IMGTYPE_GRAYA, depth, ow, oh, args)
        mx = 65535 if depth == 16 else 255
        yy, xx = np.mgrid[0:ih, 0:iw]
        img = np.empty((ih, iw, nch))
        for c in range(nch):
            g = 0.5 + 0.5 * np.sin(xx * r.uniform(0.01, 0.3) + yy * r.uniform(0.01, 0.3) + r.uniform(0, 6))
            g = np.where((xx // int(r.integers(5, 60)) + yy // int(r.integers(5, 60))) % 3 == 0, np.round(g), g)   # hard edges, exact 0 / 1
            img[..., c] = g
        img[: ih // 5, : iw // 3, nch - 1] = 0.0
        img[ih // 2:, iw // 2:, nch - 1] = 1.0
        codes = np.round(img * mx).astype(np.uint16)
        if depth == 16:
            rows = np.stack([codes >> 8, codes & 255], axis=-1).astype(np.uint8).reshape(ih, iw * nch * 2)
        else:
            rows = codes.astype(np.uint8).reshape(ih, iw * nch)
        try:
            _, b = configs.run_host(ref, cfg, rows)
        except Exception as e:  # noqa: BLE001
            print("reference refused", args, e, flush=True)
            continue
        _, a = configs.run_host(prod, cfg, rows)
        nd = int(np.count_nonzero(a != b))
        samplesa += a.size
        if nd:
            bada += 1
            print(f"DIFFALPHA seed {seed} {cfg.args!r} {iw}x{ih} depth {depth} nch {nch}: {nd} bytes", flush=True)
    print(f"images with alpha (gather-form and tile kernels, level model): {n} cases, {samplesa} output bytes compared, {bada} cases differ", flush=True)
    bad += bada
    big = 0
    for dither in ("-dither o", "-dither r", "", "-dither halftone"):
        for flt, ow, oh in (("lanczos", 3840, 2160), ("catrom", 2560, 1440), ("mitchell", 3833, 2160)):
            cfg = configs.Config("big", 7680, 4320, L.IW_IMGTYPE_RGB if flt != "mitchell" else L.IW_IMGTYPE_GRAY, 8, ow, oh,
                                 f"-width {ow} -height {oh} -filter {flt} {dither}")
            rows = configs.synth_rows(prod, cfg, seed=0xABCDEF + ow)
            _, a = configs.run_host(prod, cfg, rows)
            _, b = configs.run_host(ref, cfg, rows)
            nd = int(np.count_nonzero(a != b))
            big += nd
            print(f"8K {flt} -> {ow}x{oh} [{dither or 'no dither'}]: {nd} differing bytes of {a.size}", flush=True)
    return 1 if (bad or big) else 0


if __name__ == "__main__":
    sys.exit(main())
