"""Distinct contexts may run concurrently from different host threads (reference contract: no locks, no mutable
globals -- SURVEY section 8b "Threading"; it is how a batch of images gets parallelism).  Here every thread drives
its own contexts through iw_process_image on the same GPU; every result must equal the reference's."""
import threading

import numpy as np
import pytest

from imageworsener_b200 import configs, libimagew as L


@pytest.mark.gpu
def test_concurrent_contexts_from_several_threads(prod, ref):
    shapes = [
        configs.Config("t0", 1920, 1080, L.IW_IMGTYPE_RGB, 8, 256, 144, "-width 256 -height 144"),
        configs.Config("t1", 1024, 768, L.IW_IMGTYPE_RGB, 8, 512, 384, "-width 512 -height 384 -filter lanczos -dither o"),
        configs.Config("t2", 640, 480, L.IW_IMGTYPE_GRAY, 8, 320, 240, "-width 320 -height 240 -filter catrom -dither r"),
        configs.Config("t3", 300, 200, L.IW_IMGTYPE_RGBA, 16, 450, 300, "-width 450 -height 300 -filter mitchell -depth 16"),
    ]
    nthreads, per_thread = 6, 8
    work = []
    for t in range(nthreads):
        for k in range(per_thread):
            cfg = shapes[(t + k) % len(shapes)]
            rows = configs.synth_rows(prod, cfg, seed=0xC0FFEE + 97 * t + k)
            work.append((t, cfg, rows))
    expected = [configs.run_host(ref, cfg, rows)[1] for _, cfg, rows in work]
    got = [None] * len(work)
    errors = []

    def run(t):
        try:
            for i, (tt, cfg, rows) in enumerate(work):
                if tt == t:
                    got[i] = configs.run_host(prod, cfg, rows)[1]
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for i, (e, g) in enumerate(zip(expected, got)):
        assert g is not None and np.array_equal(e, g), f"work item {i} ({work[i][1].name}) differs"
