#!/usr/bin/env python
"""Generate tests/golden/golden_cases.{json,npz} from the reference's own regression suite.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

It reads the command lines of the reference's tests/runtest (expanding its shell loops/variables), keeps
the ones whose input is a PNG this fixture reader models exactly and whose output is PNG / PNM / PAM,
decodes the input the way the reference's PNG reader would (pngmini.read_png_as_iw_input) and the expected
output from tests/expected into canonical RGBA, and stores both.  It then replays every case through
oracle/_ref (the unmodified reference build) via the same harness the tests use and REFUSES to write a case
that the reference itself does not reproduce -- so every committed vector is known to be reachable through
the public API with the decoded input.
"""
from __future__ import annotations

import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import harness  # noqa: E402
import pngmini  # noqa: E402  (tests/golden is on sys.path because this script lives there)
from imageworsener_b200 import libimagew as L  # noqa: E402
import oracle  # noqa: E402  (test infrastructure: the reference build / port)
from imageworsener_b200.jobs import Job  # noqa: E402

REF = os.environ.get("IW_REFERENCE", "/root/reference")
TESTS = os.path.join(REF, "tests")


def expand_runtest():
    """Yield (src, dst, argstring) for every `$IW ...` line of tests/runtest."""
    text = open(os.path.join(TESTS, "runtest")).read().replace("\\\n", " ")
    var = {"SCALE": "-width 35 -height 35", "SCALE2": "-width 24 -height 24", "SMALL": "-width 15 -height 15",
           "CMPR": "-zipcmprlevel 9", "DCMPR": "-zipcmprlevel -1"}
    lines = text.splitlines()
    out = []

    def subst(s, env):
        def rep(m):
            k = m.group(1) or m.group(2)
            return env.get(k, m.group(0))
        return re.sub(r"\$\{(\w+)\}|\$(\w+)", rep, s)

    i = 0
    stack = []   # (varname, values, body_start)
    env = dict(var)

    def emit(line, env):
        line = subst(line.strip(), env)
        toks = line.replace('"', "").split()
        if not toks or toks[0] != "$IW":
            return
        src, dst = toks[1], toks[2]
        out.append((src, dst, " ".join(toks[3:])))

    while i < len(lines):
        ln = lines[i].strip()
        m = re.match(r"for (\w+) in (.*)$", ln)
        if m:
            name, vals = m.group(1), m.group(2).split()
            body = []
            i += 1
            assert lines[i].strip() == "do"
            i += 1
            while lines[i].strip() != "done":
                body.append(lines[i])
                i += 1
            for v in vals:
                e = dict(env)
                e[name] = v
                for b in body:
                    emit(b, e)
        elif ln.startswith("$IW"):
            emit(ln, env)
        i += 1
    return out


def main():
    ref = oracle.load_reference()
    cases_meta = []
    blob = {}
    skipped = []
    n_try = 0
    for src, dst, args in expand_runtest():
        name = os.path.splitext(os.path.basename(dst))[0] + "_" + os.path.splitext(dst)[1][1:]
        ext = os.path.splitext(dst)[1].lower()
        if not src.endswith(".png") or ext not in (".png", ".ppm", ".pgm", ".pnm", ".pam"):
            skipped.append((dst, "format outside the fixture reader"))
            continue
        srcpath = os.path.join(TESTS, src.replace("actual/", "expected/"))
        exppath = os.path.join(TESTS, "expected", os.path.basename(dst))
        if not os.path.exists(srcpath) or not os.path.exists(exppath):
            skipped.append((dst, "file missing"))
            continue
        inp = pngmini.read_png_as_iw_input(srcpath)
        if inp["sbit"]:
            skipped.append((dst, "sBIT input"))
            continue
        if inp["phys"] is not None and inp["phys"][0] != inp["phys"][1]:
            skipped.append((dst, "non-square pixel density changes CLI size fitting"))
            continue
        if "-density" in args:
            skipped.append((dst, "density option (codec metadata)"))
            continue
        outfmt = {".png": "png", ".ppm": "ppm", ".pgm": "pgm", ".pnm": "pnm", ".pam": "pam"}[ext]
        try:
            job = Job.parse(args, outfmt=outfmt)
        except Exception as e:  # noqa: BLE001
            skipped.append((dst, f"option parse: {e}"))
            continue
        case = {"name": name, "src": src, "dst": dst, "args": args, "outfmt": outfmt}
        if ext == ".png":
            try:
                depth, exp, epng = pngmini.png_to_canonical(exppath)
            except NotImplementedError as e:
                skipped.append((dst, str(e)))
                continue
            if epng.sbit is not None:
                skipped.append((dst, "sBIT output (reduced depth scaled by the PNG writer)"))
                continue
            case.update(exp_kind="png", exp_depth=depth, exp_maxval=(1 << depth) - 1, exp_alpha_missing=False)
        else:
            mx, exp = pngmini.pnm_to_canonical(exppath)
            head = open(exppath, "rb").read(2)
            nch = 0
            if head == b"P7":
                hdr = open(exppath, "rb").read(200).decode("latin1")
                nch = int(re.search(r"DEPTH (\d+)", hdr).group(1))
            case.update(exp_kind="pnm", exp_depth=8 if mx < 256 else 16, exp_maxval=mx,
                        exp_alpha_missing=(head != b"P7" or nch in (1, 3)))
        case["input"] = {k: inp[k] for k in ("imgtype", "bit_depth", "width", "height", "native_grayscale",
                                              "cs", "bkgd_label")}
        full = dict(case)
        full["input"] = dict(case["input"], rows=inp["rows"])
        full["expected"] = exp.astype(np.uint32)
        n_try += 1
        # replay through the unmodified reference; default optimisations OFF, as the tests will run it
        job = Job.parse(args, outfmt=outfmt, disable_all_opt=True)
        res = harness.run_job(ref, full["input"], job)
        if not res["ok"]:
            skipped.append((dst, "reference failed: " + str(res["err"])))
            continue
        nd = harness.compare_with_golden(res, full)
        if nd != 0:
            skipped.append((dst, f"reference replay differs from tests/expected in {nd} samples "
                                 "(harness cannot model this case)"))
            continue
        cases_meta.append(case)
        blob[name + "__in"] = inp["rows"]
        blob[name + "__exp"] = exp.astype(np.uint16)
    json.dump({"source": "jsummers/imageworsener tests/runtest + tests/srcimg + tests/expected (v1.3.5)",
               "cases": cases_meta, "skipped": skipped},
              open(os.path.join(HERE, "golden_cases.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_cases.npz"), **blob)
    print(f"{len(cases_meta)} cases written ({n_try} attempted); {len(skipped)} skipped")
    for s in skipped:
        print("  skipped", s)


if __name__ == "__main__":
    main()
