"""On-box check of the second-generation kernels: byte parity against the reference build (reduced cfg2 and a few
odd shapes) and against the generation-1 kernels at full size, then device timings for option variants.

    python tools/check_gen2.py [--variants "cuda:strip2=96 cuda:seg2=64" "cuda:strip2=192"]
"""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imageworsener_b200 import configs, libimagew as L  # noqa: E402
import oracle  # noqa: E402  (test infrastructure: the reference build / port)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", nargs="*", default=[""])
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--quick", action="store_true", help="only the cfg2 shapes (IWB_QUICK builds)")
    args = ap.parse_args()
    prod = L.load_product()
    ok = True
    if not args.skip_parity:
        ref = oracle.load_reference()
        shapes = [("cfg2", 16), ("cfg2", 4)] if args.quick else [("cfg2", 16), ("cfg2", 4), ("cfg4box", 1), ("cfg4auto", 2)]
        for name, scale in shapes:
            cfg = configs.get(name, scale)
            rows = configs.synth_rows(prod, cfg)
            _, b = configs.run_host(ref, cfg, rows)
            for gen in (1, 2):
                _, a = configs.run_host(prod, cfg, rows, options={"cuda:gen": gen})
                nd = int(np.count_nonzero(a != b))
                print(f"parity {name}/{scale} gen{gen}: {nd} differing bytes of {a.size}", flush=True)
                ok = ok and nd == 0
        # extra shapes: odd sizes, gray, no dither, random dither
        extra = [
            configs.Config("odd", 1000, 700, L.IW_IMGTYPE_RGB, 8, 617, 333, "-width 617 -height 333 -filter lanczos -dither o"),
            configs.Config("gray", 1024, 768, L.IW_IMGTYPE_GRAY, 8, 400, 300, "-width 400 -height 300 -filter mitchell"),
            configs.Config("rand", 1024, 768, L.IW_IMGTYPE_RGB, 8, 512, 384, "-width 512 -height 384 -filter catrom -dither r"),
            configs.Config("up", 320, 240, L.IW_IMGTYPE_RGB, 8, 512, 400, "-width 512 -height 400 -filter hermite -dither halftone"),
            configs.Config("wide", 4099, 64, L.IW_IMGTYPE_RGB, 8, 2011, 37, "-width 2011 -height 37 -filter lanczos2"),
        ]
        for cfg in ([] if args.quick else extra):
            rows = configs.synth_rows(prod, cfg)
            _, b = configs.run_host(ref, cfg, rows)
            for gen in (1, 2):
                _, a = configs.run_host(prod, cfg, rows, options={"cuda:gen": gen})
                nd = int(np.count_nonzero(a != b))
                print(f"parity {cfg.name} gen{gen}: {nd} differing bytes of {a.size}", flush=True)
                ok = ok and nd == 0
        cfg = configs.get("cfg2")
        rows = configs.synth_rows(prod, cfg)
        _, a1 = configs.run_host(prod, cfg, rows, options={"cuda:gen": 1})
        _, a2 = configs.run_host(prod, cfg, rows, options={"cuda:gen": 2})
        nd = int(np.count_nonzero(a1 != a2))
        print(f"parity cfg2 full gen2 vs gen1: {nd} differing bytes, fnv {configs.fnv1a64(prod, a2):016x}", flush=True)
        ok = ok and nd == 0
    for v in args.variants:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "1"]
        env = dict(os.environ)
        for o in v.split():
            if o.startswith("lib="):      # a tuning build under imageworsener_b200/_variants
                env["IWB200_LIB"] = os.path.join(ROOT, "imageworsener_b200", "_variants", o[4:] + ".so")
            else:
                cmd += ["--opt", o]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            k = d["roofline"]["kernel_ms_per_step"]
            print(f"timing [{v}]: step {d['ms_per_step']:.4f} ms  " + "  ".join(f"{n} {x['ms_per_step']:.4f}" for n, x in k.items()), flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"timing [{v}]: failed: {e}\n{r.stdout[-500:]}\n{r.stderr[-1500:]}", flush=True)
    print("PARITY", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
