"""The BASELINE.json configurations.

  * reduced sizes: CUDA output == reference output, byte for byte (the reference finishes these in seconds);
  * full sizes: FNV-1a-64 fingerprint of the CUDA output == the fingerprint recorded from the reference build by
    tests/golden/make_known_answers.py (known_answers.json);
  * size-independent property: processing the image as N output-row bands (the multi-GPU sharding unit) and
    concatenating them gives the same bytes as processing it whole -- ordered-dither phase, checkerboard
    phase and the random-dither draw index must all use global coordinates for this to hold.
"""
import json
import os

import numpy as np
import pytest

from imageworsener_b200 import configs

KNOWN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))


def _bands(lib, cfg, rows, n):
    parts = []
    for i in range(n):
        img, px = configs.run_host(lib, cfg, rows, options={"cuda:band": f"{i}/{n}"})
        parts.append(px)
    return np.concatenate(parts, axis=0)


@pytest.mark.parametrize("name,scale", [("cfg2", 16), ("cfg3", 32), ("cfg4auto", 4), ("cfg4box", 4), ("cfg5", 128)])
def test_host_logic_configs_match_reference(emu, ref, name, scale):
    cfg = configs.get(name, scale)
    rows = configs.synth_rows(emu, cfg)
    _, a = configs.run_host(emu, cfg, rows)
    _, b = configs.run_host(ref, cfg, rows)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name,scale,n", [("cfg2", 16, 3), ("cfg5", 128, 4), ("cfg3", 32, 2)])
def test_host_logic_bands_concatenate_to_whole(emu, name, scale, n):
    cfg = configs.get(name, scale)
    rows = configs.synth_rows(emu, cfg)
    _, whole = configs.run_host(emu, cfg, rows)
    assert np.array_equal(_bands(emu, cfg, rows, n), whole)


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale", [("cfg2", 4), ("cfg3", 8), ("cfg4auto", 1), ("cfg4box", 1), ("cfg5", 16)])
def test_cuda_configs_match_reference(prod, ref, name, scale):
    cfg = configs.get(name, scale)
    rows = configs.synth_rows(prod, cfg)
    ia, a = configs.run_host(prod, cfg, rows)
    ib, b = configs.run_host(ref, cfg, rows)
    assert (ia.imgtype, ia.bit_depth, ia.width, ia.height, ia.bpr) == (ib.imgtype, ib.bit_depth, ib.width, ib.height, ib.bpr)
    nd = int(np.count_nonzero(a != b))
    assert nd == 0, f"{nd} differing bytes of {a.size}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg4auto", "cfg4box", "cfg2", "cfg3", "cfg5"])
def test_cuda_full_size_known_answer(prod, name):
    if name not in KNOWN:
        pytest.skip("no known answer recorded for " + name)
    cfg = configs.get(name)
    rows = configs.synth_rows(prod, cfg)
    assert "%016x" % configs.fnv1a64(prod, rows) == KNOWN[name]["input_fnv1a64"]
    img, px = configs.run_host(prod, cfg, rows)
    assert (img.width, img.height, img.bit_depth, img.imgtype) == (
        KNOWN[name]["out_w"], KNOWN[name]["out_h"], KNOWN[name]["bit_depth"], KNOWN[name]["imgtype"])
    assert "%016x" % configs.fnv1a64(prod, px) == KNOWN[name]["fnv1a64"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale,n", [("cfg2", 2, 8), ("cfg5", 8, 5), ("cfg3", 4, 3)])
def test_cuda_bands_concatenate_to_whole(prod, name, scale, n):
    cfg = configs.get(name, scale)
    rows = configs.synth_rows(prod, cfg)
    _, whole = configs.run_host(prod, cfg, rows)
    assert np.array_equal(_bands(prod, cfg, rows, n), whole)
