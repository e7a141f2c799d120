"""Differential tests: the same seeded random image + command line through this library and through the
unmodified reference build must give identical output bytes (same type, depth, size, every byte).

not-gpu: host-emulation build (host logic);  gpu: the CUDA library through the C ABI.
"""
import numpy as np
import pytest

import harness
import randjobs
from imageworsener_b200.jobs import Job


def _one(lib, ref, seed, maker=randjobs.make, **kw):
    inp, args, outfmt = maker(seed, **kw)
    ja = Job.parse(args, outfmt=outfmt, disable_all_opt=True)
    jb = Job.parse(args, outfmt=outfmt, disable_all_opt=True)
    a = harness.run_job(lib, inp, ja)
    b = harness.run_job(ref, inp, jb)
    tag = f"seed={seed} args={args!r} outfmt={outfmt} in={inp['imgtype']:#x}/{inp['bit_depth']} {inp['width']}x{inp['height']}"
    assert a["ok"] == b["ok"], (tag, a["err"], b["err"])
    if not a["ok"]:
        return
    for k in ("imgtype", "bit_depth", "sampletype", "width", "height", "bpr", "reduced_maxcolors", "out_cstype"):
        assert a[k] == b[k], (tag, k, a[k], b[k])
    nd = int(np.count_nonzero(a["pixels"] != b["pixels"]))
    assert nd == 0, (tag, f"{nd} differing bytes of {a['pixels'].size}")
    assert a["has_bkgdlabel"] == b["has_bkgdlabel"], tag
    if a["has_bkgdlabel"]:
        assert a["bkgdlabel"] == b["bkgdlabel"], tag


@pytest.mark.parametrize("seed", range(300))
def test_host_logic_matches_reference(emu, ref, seed):
    _one(emu, ref, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(1000, 1400))
def test_cuda_matches_reference_small(prod, ref, seed):
    _one(prod, ref, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(2000, 2060))
def test_cuda_matches_reference_medium(prod, ref, seed):
    _one(prod, ref, seed, max_in=700, max_out=900, allow_errdiff=False)


# floating-point input samples (reference: get_raw_sample_flt32 + x_to_linear_sample per sample,
# src/imagew-main.c:53-66, :129-145, :348-372), in linear and non-linear colourspaces, with and without alpha
@pytest.mark.parametrize("seed", range(5000, 5060))
def test_host_logic_float_input(emu, ref, seed):
    _one(emu, ref, seed, maker=randjobs.make_float)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(5100, 5180))
def test_cuda_float_input(prod, ref, seed):
    _one(prod, ref, seed, maker=randjobs.make_float)
