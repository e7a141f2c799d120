"""N3 of SURVEY section 8(f): the reference's command-line tool and its dependency-free file-format modules
(src/imagew-cmd.c, src/imagew-{pnm,bmp,miff,tiff,gif,allfmts}.c), compiled UNMODIFIED from the reference sources and
linked against this repo's library instead of the reference core (oracle/Makefile, target `cli`).  Every command line of
the reference's own regression script (tests/runtest, as recorded in tests/golden/golden_cases.json) is run through
that tool and through the all-reference build of the same tool; the files they write must be identical byte for byte
-- option parsing, the API calls the tool makes, the pipeline, the post-pass optimiser (palette / gray / depth choices
show up in BMP and PAM headers) and the utility functions the codecs call are all on that path.

The PNG / JPEG / WebP modules need third-party libraries that are not in this image, so inputs are handed over as PAM
files (written from the decoded fixtures) and outputs alternate between PAM, BMP, MIFF and TIFF.
"""
import os
import subprocess

import numpy as np
import pytest

import harness
from imageworsener_b200 import libimagew as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "imagew_ref")
OUR_CLI = os.path.join(ROOT, "oracle", "_ref", "imagew_b200")
EMU_CLI = os.path.join(ROOT, "oracle", "_ref", "imagew_emu")
CASES = harness.load_golden()
FORMATS = ["pam", "bmp", "miff", "tif"]


def write_pam(path, inp):
    """The fixture's decoded input (packed rows at 1-16 bits) as a PAM file the reference's PNM module reads."""
    h, w, depth = inp["height"], inp["width"], inp["bit_depth"]
    nch = L.imgtype_num_channels(inp["imgtype"])
    rows = inp["rows"]
    if depth < 8:
        bits = np.unpackbits(rows, axis=1)[:, : w * nch * depth].reshape(h, w * nch, depth)
        codes = np.zeros((h, w * nch), np.uint16)
        for k in range(depth):
            codes = (codes << 1) | bits[:, :, k]
        data = codes.astype(np.uint8).tobytes()
    elif depth == 8:
        data = np.ascontiguousarray(rows[:, : w * nch]).tobytes()
    else:
        data = np.ascontiguousarray(rows[:, : w * nch * 2]).tobytes()      # already big-endian
    tupl = {1: "GRAYSCALE", 2: "GRAYSCALE_ALPHA", 3: "RGB", 4: "RGB_ALPHA"}[nch]
    with open(path, "wb") as fh:
        fh.write(f"P7\nWIDTH {w}\nHEIGHT {h}\nDEPTH {nch}\nMAXVAL {(1 << depth) - 1}\nTUPLTYPE {tupl}\nENDHDR\n".encode())
        fh.write(data)


def run_cli(binary, src, dst, args):
    cmd = [binary, src, dst] + args.split() + ["-quiet"]
    r = subprocess.run(cmd, capture_output=True, timeout=120)
    data = open(dst, "rb").read() if os.path.exists(dst) else None
    return r.returncode, data, r.stderr.decode("utf-8", "replace")


def _one(binary, case, k, tmp_path):
    if not (os.path.exists(binary) and os.path.exists(REF_CLI)):
        pytest.skip("command-line tools not built (make -C oracle cli needs /root/reference)")
    src = str(tmp_path / "in.pam")
    write_pam(src, case["input"])
    fmt = FORMATS[k % len(FORMATS)]
    args = case["args"]
    rc_a, a, err_a = run_cli(binary, src, str(tmp_path / f"ours.{fmt}"), args)
    rc_b, b, err_b = run_cli(REF_CLI, src, str(tmp_path / f"ref.{fmt}"), args)
    tag = f"{case['name']} -> {fmt}: {args}"
    assert (rc_a == 0) == (rc_b == 0), (tag, err_a[-300:], err_b[-300:])
    if rc_b != 0:
        return
    assert a is not None and b is not None, tag
    if a != b and fmt == "tif" and len(a) == len(b):
        # The reference's TIFF writer emits a colour map of 2^bits entries but its optimiser only fills the entries in
        # use and does not clear the rest (src/imagew-opt.c:1036, iw_malloc): the tail of the map is whatever the heap
        # held.  This library clears the palette, so only a few map bytes may differ -- nothing else.
        nd = sum(1 for x, y in zip(a, b) if x != y)
        assert nd <= 96, f"{tag}: {nd} differing bytes"
        return
    assert a == b, f"{tag}: files differ ({len(a)} vs {len(b)} bytes)"


@pytest.mark.parametrize("k", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_reference_tool_on_host_emulation_writes_identical_files(k, tmp_path):
    _one(EMU_CLI, CASES[k], k, tmp_path)


# (every process pays for a CUDA context: every second command line on the device, all of them on the emulation above)
@pytest.mark.gpu
@pytest.mark.parametrize("k", range(0, len(CASES), 2), ids=[c["name"] for c in CASES[::2]])
def test_reference_tool_on_cuda_library_writes_identical_files(k, tmp_path):
    _one(OUR_CLI, CASES[k], k, tmp_path)
