"""cuda:devices -- one iw_process_image() call spread over several GPUs by the library itself (csrc/pipeline.cu:
iwb200_run_host_multi): output-row bands, one host thread + stream set per listed device, every band copied device ->
host straight into the caller's output buffer.  The same ordinal may be listed more than once, which lets a one-GPU box
exercise the whole mechanism (threads, per-band jobs, absolute dither / PRNG coordinates); the tests with two different
devices run where the box has them.  Reference semantics of what must be preserved across bands: ordered dither and
checkerboard use absolute (x, y) (src/imagew-main.c:457-484, :1082-1085), random dither one draw per sample in raster
order (:719-723, :1002-1012)."""
import numpy as np
import pytest

import harness
import randjobs
from imageworsener_b200 import configs, libimagew as L
from imageworsener_b200.jobs import Job

RGB, GRAY, RGBA = L.IW_IMGTYPE_RGB, L.IW_IMGTYPE_GRAY, L.IW_IMGTYPE_RGBA

SHAPES = [
    ("lanczos_o", 1024, 768, RGB, 8, 512, 384, "-filter lanczos -dither o"),
    ("catrom_r_gray", 1000, 700, GRAY, 8, 333, 233, "-filter catrom -grayscale -dither r"),
    ("rgba16_bkgd", 300, 260, RGBA, 16, 450, 390, "-filter mitchell -depth 16 -bkgd e42,0f5 -checkersize 7"),
    ("gray16_r", 640, 400, GRAY, 16, 320, 200, "-filter catrom -grayscale -dither r"),
    ("up_o", 320, 240, RGB, 8, 512, 400, "-filter hermite -dither halftone"),
]


def _cfg(shape):
    name, iw, ih, typ, depth, ow, oh, extra = shape
    return configs.Config(name, iw, ih, typ, depth, ow, oh, f"-width {ow} -height {oh} {extra}")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0,0"])
def test_bands_on_listed_devices_equal_the_reference(prod, ref, shape, devices):
    cfg = _cfg(shape)
    rows = configs.synth_rows(prod, cfg)
    ia, a = configs.run_host(prod, cfg, rows, options={"cuda:devices": devices})
    ib, b = configs.run_host(ref, cfg, rows)
    assert (ia.width, ia.height, ia.bit_depth, ia.imgtype, ia.bpr) == (ib.width, ib.height, ib.bit_depth, ib.imgtype, ib.bpr)
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_two_different_devices_equal_one_device(prod):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = configs.get("cfg2", scale=2)
    rows = configs.synth_rows(prod, cfg)
    _, a = configs.run_host(prod, cfg, rows, options={"cuda:devices": "0,1"})
    _, b = configs.run_host(prod, cfg, rows, options={"cuda:device": "1"})
    _, c = configs.run_host(prod, cfg, rows)
    assert np.array_equal(a, c) and np.array_equal(b, c)


@pytest.mark.gpu
def test_error_diffusion_refuses_cuda_devices(prod):
    cfg = configs.Config("ed", 200, 160, RGB, 8, 100, 80, "-width 100 -height 80 -dither f")
    rows = configs.synth_rows(prod, cfg)
    with pytest.raises(L.IwError, match="error-diffusion"):
        configs.run_host(prod, cfg, rows, options={"cuda:devices": "0,0"})


@pytest.mark.gpu
@pytest.mark.parametrize("bad", ["a,b", "0,-1", "0;1", ",", "0,1x"])
def test_malformed_device_lists_fail_loudly(prod, bad):
    cfg = configs.Config("bad", 64, 64, RGB, 8, 32, 32, "-width 32 -height 32")
    rows = configs.synth_rows(prod, cfg)
    with pytest.raises(L.IwError):
        configs.run_host(prod, cfg, rows, options={"cuda:devices": bad})


@pytest.mark.gpu
def test_unknown_device_ordinal_fails_loudly(prod):
    cfg = configs.Config("bad", 256, 256, RGB, 8, 128, 128, "-width 128 -height 128")
    rows = configs.synth_rows(prod, cfg)
    with pytest.raises(L.IwError, match="device"):
        configs.run_host(prod, cfg, rows, options={"cuda:devices": "0,512"})


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(7000, 7040))
def test_random_jobs_in_two_bands(prod, ref, seed):
    """Seeded random jobs (every filter / edge policy / depth / alpha / background, no error diffusion) through two bands."""
    inp, args, outfmt = randjobs.make(seed, max_in=300, max_out=400, allow_errdiff=False)
    a = harness.run_job(prod, inp, Job.parse(args, outfmt=outfmt, disable_all_opt=True), backend_opts={"cuda:devices": "0,0"})
    b = harness.run_job(ref, inp, Job.parse(args, outfmt=outfmt, disable_all_opt=True))
    assert a["ok"] == b["ok"], (args, a["err"], b["err"])
    if a["ok"]:
        assert np.array_equal(a["pixels"], b["pixels"]), args


@pytest.mark.parametrize("seed", range(7100, 7130))
def test_host_logic_two_bands(emu, ref, seed):
    """The band split, the per-band jobs and the threads on the GPU-less box (test-only host emulation of the kernels)."""
    inp, args, outfmt = randjobs.make(seed, allow_errdiff=False)
    ctx_opts = {"cuda:devices": "0,0"}
    a = harness.run_job(emu, inp, Job.parse(args, outfmt=outfmt, disable_all_opt=True), backend_opts=ctx_opts)
    b = harness.run_job(ref, inp, Job.parse(args, outfmt=outfmt, disable_all_opt=True))
    assert a["ok"] == b["ok"], (args, a["err"], b["err"])
    if a["ok"]:
        assert np.array_equal(a["pixels"], b["pixels"]), args
