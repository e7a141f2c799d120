#!/usr/bin/env python
"""Full-size known answers for the BASELINE.json configurations.

Runs the UNMODIFIED reference build (oracle/_ref) on the deterministic xorshift64 noise input of each
configuration (imageworsener_b200/configs.py) and records the FNV-1a-64 fingerprint of its output buffer in
tests/golden/known_answers.json.  The GPU parity tests compare the CUDA path's fingerprint at the full
BASELINE sizes against these, where re-running the reference would take minutes (cfg5: ~5 min, 6 GiB).

Build container only:  python tests/golden/make_known_answers.py [cfg ...]
(The fingerprints printed in SURVEY.md section 8(c) came from a throw-away probe whose exact input layout was not
recorded and could not be reproduced; these replace them and are reproducible with this script.)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from imageworsener_b200 import _build, configs, libimagew as L  # noqa: E402
import oracle  # noqa: E402  (test infrastructure: the reference build / port)

OUT = os.path.join(HERE, "known_answers.json")


def main():
    names = sys.argv[1:] or ["cfg4auto", "cfg4box", "cfg2", "cfg3", "cfg5"]
    helper = L.Lib(_build.build_emulator())   # only for the synth/hash helpers (plain host C)
    ref = oracle.load_reference()
    known = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        cfg = configs.get(name)
        rows = configs.synth_rows(helper, cfg)
        t = time.time()
        img, px = configs.run_host(ref, cfg, rows)
        dt = time.time() - t
        known[cfg.name] = {"fnv1a64": "%016x" % configs.fnv1a64(helper, px), "out_w": img.width, "out_h": img.height,
                           "imgtype": img.imgtype, "bit_depth": img.bit_depth, "bpr": img.bpr,
                           "input_fnv1a64": "%016x" % configs.fnv1a64(helper, rows),
                           "reference_seconds_1thread": round(dt, 2), "args": cfg.args}
        print(cfg.name, known[cfg.name], flush=True)
        json.dump(known, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
