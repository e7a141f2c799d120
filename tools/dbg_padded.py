import sys, numpy as np
sys.path.insert(0, '/root/repo')
from imageworsener_b200 import configs, libimagew as L
prod = L.load_product(); ref = L.load_reference()
def run(w, h, ow, oh, flt, opts, typ=L.IW_IMGTYPE_RGB, pad=9):
    nch = L.imgtype_num_channels(typ)
    cfg = configs.Config("padded", w, h, typ, 8, ow, oh, f"-width {ow} -height {oh} -filter {flt} -dither o")
    tight = configs.synth_rows(prod, cfg)
    bpr = w * nch + pad
    rows = np.zeros((h, bpr), np.uint8); rows[:, : w * nch] = tight; rows[:, w * nch:] = 0xA5
    _, b = configs.run_host(ref, cfg, tight)
    _, a = configs.run_host(prod, cfg, rows, options=opts)
    bad = np.argwhere(a != b)
    print(w, h, ow, oh, flt, opts, len(bad), "differing; px", sorted(set(bad[:, 1] // nch))[:12], "rows", sorted(set(bad[:, 0]))[:8], flush=True)
for seg in (0, 20, 40, 84, 164):
    run(333, 300, 166, 150, "lanczos", {"cuda:seg2": seg} if seg else {})
run(333, 300, 166, 300, "lanczos", {})            # horizontal only
run(333, 300, 333, 150, "lanczos", {})            # vertical only
run(335, 300, 166, 150, "lanczos", {}, pad=3)     # another ragged factor
run(333, 300, 166, 150, "catrom", {})
run(1000, 300, 498, 150, "lanczos", {}, pad=8)
