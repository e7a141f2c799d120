"""The N>1 path on CPU: two gloo ranks each process one output-row band (through the test-only host-emulation
build, since this box has no GPU), rank 0 gathers the bands and the result must equal the whole image byte for
byte.  Also pins the band planner against the library's own split and halo computation."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from imageworsener_b200 import bands, configs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, scale, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from imageworsener_b200 import _build, libimagew as L
        lib = L.Lib(_build.EMU_SO)
        cfg = configs.get(name, scale)
        rows = configs.synth_rows(lib, cfg)
        local = bands.process_band(lib, cfg, rows, rank, world)
        whole = bands.gather_bands(local, cfg.out_h, rank, world)
        if rank == 0:
            _, ref = configs.run_host(lib, cfg, rows)
            q.put(bool(np.array_equal(whole, ref)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,scale", [("cfg2", 16), ("cfg5", 128)])
def test_two_rank_band_sharding_gloo(emu, name, scale):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, scale, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_band_planner_matches_library(emu):
    lib = emu
    cfg = configs.get("cfg2", 16)
    rows = configs.synth_rows(lib, cfg)
    for n in (1, 2, 3, 7):
        tot = 0
        for i in range(n):
            px = bands.process_band(lib, cfg, rows, i, n)
            r0, r1 = bands.band_rows(cfg.out_h, i, n)
            assert px.shape[0] == r1 - r0
            tot += px.shape[0]
        assert tot == cfg.out_h


def test_band_halo_is_filter_support(emu):
    """A band needs its own input rows plus a halo of about radius*reduction rows (Lanczos3, 2x down: 6)."""
    lib = emu
    dll = lib.dll
    dll.iwb200_job_create.restype = C.c_void_p
    dll.iwb200_job_create.argtypes = [C.c_void_p, C.c_int]
    dll.iwb200_job_destroy.argtypes = [C.c_void_p]
    dll.iwb200_job_band_input_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    cfg = configs.get("cfg2", 16)
    ctx = lib.iw_create_context(None)
    lib.set_input_from_numpy(ctx, configs.synth_rows(lib, cfg), cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    cfg.apply(lib, ctx)
    job = dll.iwb200_job_create(ctx, 0)
    assert job
    a, b = C.c_int(), C.c_int()
    assert dll.iwb200_job_band_input_rows(job, 40, 80, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (2 * 40 - 5, 2 * 79 + 6)
    assert dll.iwb200_job_band_input_rows(job, 0, cfg.out_h, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (0, cfg.in_h - 1)
    dll.iwb200_job_destroy(job)
    lib.iw_destroy_context(ctx)
