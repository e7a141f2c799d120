"""Where the end-to-end time of iw_process_image goes (plan, H2D span, kernel span after the last H2D, D2H span after the
last kernel): python tools/e2e_breakdown.py CONFIG [option=value ...]"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from imageworsener_b200 import configs, libimagew as L
prod = L.load_product()
dll = prod.dll
dll.iwb200_get_timings.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 4
name = sys.argv[1]
opts = dict(a.split("=", 1) for a in sys.argv[2:])
cfg = configs.get(name)
rows = configs.synth_rows(prod, cfg)
tt = []
for it in range(8):
    ctx = prod.iw_create_context(None)
    prod.set_input_from_numpy(ctx, rows, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    for k, v in opts.items():
        prod.iw_set_option(ctx, k.encode(), v.encode())
    cfg.apply(prod, ctx)
    t0 = time.perf_counter()
    ok = prod.iw_process_image(ctx)
    dt = (time.perf_counter() - t0) * 1e3
    a = [C.c_double() for _ in range(4)]
    dll.iwb200_get_timings(ctx, *[C.byref(x) for x in a])
    tt.append((dt, [round(x.value, 3) for x in a]))
    prod.iw_destroy_context(ctx)
print(name, opts, [(round(d, 3), a) for d, a in tt[4:]])
