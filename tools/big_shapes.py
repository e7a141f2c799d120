"""Full-size parity runs of the newer fast-path variants (enlargements, gray as RGB, holes) against the reference build."""
import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageworsener_b200 import configs, libimagew as L
import oracle  # noqa: E402  (test infrastructure: the reference build / port)
prod, ref = L.load_product(), oracle.load_reference()
RGB, GRAY = L.IW_IMGTYPE_RGB, L.IW_IMGTYPE_GRAY
bad = 0
for name, iw, ih, typ, depth, ow, oh, extra, opts in [
    ("up4_mitchell", 1920, 1080, RGB, 8, 7680, 4320, "-filter mitchell -dither o", {}),
    ("up2_lanczos_chunks", 1920, 1080, RGB, 8, 3840, 2160, "-filter lanczos -dither r", {"cuda:chunk_mb": 16}),
    ("up3_catrom_gray_rgb", 2000, 1000, GRAY, 8, 6000, 3000, "-filter catrom -dither r", {}),
    ("up2p5_gray16", 1600, 900, GRAY, 16, 4000, 2250, "-filter lanczos2 -grayscale -dither o", {"cuda:chunk_mb": 8}),
    ("down3_catrom_big_bands", 7680, 4320, RGB, 8, 2560, 1440, "-filter catrom -dither o", {"cuda:chunk_mb": 32}),
]:
    cfg = configs.Config(name, iw, ih, typ, depth, ow, oh, f"-width {ow} -height {oh} {extra}")
    rows = configs.synth_rows(prod, cfg)
    _, a = configs.run_host(prod, cfg, rows, options=opts)
    _, b = configs.run_host(ref, cfg, rows)
    nd = int(np.count_nonzero(a != b))
    bad += nd
    print(f"{name}: {nd} differing bytes of {a.size}", flush=True)
sys.exit(1 if bad else 0)
