"""Time error-diffusion dithering (cfg2 geometry, Floyd-Steinberg) on the device and on one host core.

  python tools/errdiff_probe.py [scale ...]     scale = linear shrink of 7680x4320 -> 3840x2160 (default: 8 4)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from imageworsener_b200 import configs, libimagew as L  # noqa: E402
import oracle  # noqa: E402  (CPU baseline beside the GPU number)


def cfg_errdiff(scale, dither="f"):
    k = scale
    return configs.Config("errdiff", 7680 // k, 4320 // k, L.IW_IMGTYPE_RGB, 8, 3840 // k, 2160 // k,
                          f"-width {3840 // k} -height {2160 // k} -filter lanczos -dither {dither}",
                          description=f"RGB8 2x downscale, Lanczos3, error diffusion ({dither})")


def main():
    scales = [int(a) for a in sys.argv[1:]] or [8, 4]
    prod = L.load_product()
    ref = oracle.load_reference() if os.path.exists(oracle.REFERENCE_SO) else None
    out = []
    for k in scales:
        cfg = cfg_errdiff(k)
        rows = oracle.synth_xorshift(cfg.in_bytes, configs.XORSHIFT_SEED).reshape(cfg.in_h, cfg.in_bpr)
        configs.run_host(prod, cfg, rows)                     # warm (plan cache, pools)
        t0 = time.perf_counter()
        img, px = configs.run_host(prod, cfg, rows)
        t_gpu = time.perf_counter() - t0
        rec = {"scale": k, "in": [cfg.in_w, cfg.in_h], "out": [cfg.out_w, cfg.out_h], "gpu_e2e_s": t_gpu}
        if ref is not None:
            t0 = time.perf_counter()
            rimg, rpx = configs.run_host(ref, cfg, rows)
            rec["cpu_1thread_s"] = time.perf_counter() - t0
            rec["differing_bytes"] = int(np.count_nonzero(px != rpx))
        print(json.dumps(rec), flush=True)
        out.append(rec)


if __name__ == "__main__":
    main()
