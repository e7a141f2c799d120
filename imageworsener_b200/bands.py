"""Output-row band sharding: the unit of multi-GPU parallelism for one large image.

The vertical pass runs first, so an output-row band [r0, r1) depends only on the input rows its vertical filter
windows touch (band + a halo of filter-support rows); both passes are then band-local.  Dither phase,
checkerboard phase and the random-dither draw index are functions of ABSOLUTE output coordinates inside the
library, so concatenated bands are byte-identical to the whole image (tests/test_configs.py,
tests/test_multigpu.py).  There is no collective inside the kernels; the only exchange is the final gather of
the output bands to rank 0 (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).

Error-diffusion dither does not shard (every row depends on the previous one): replicas only.
"""
from __future__ import annotations

import numpy as np


def band_rows(out_h: int, index: int, count: int) -> tuple[int, int]:
    """Same split as the library's "cuda:band" option (csrc/process.cpp)."""
    r0 = (out_h * index) // count
    r1 = (out_h * (index + 1)) // count
    if r1 <= r0:
        r1 = min(r0 + 1, out_h)
    return r0, r1


def all_bands(out_h: int, count: int) -> list[tuple[int, int]]:
    return [band_rows(out_h, i, count) for i in range(count)]


def process_band(lib, cfg, rows: np.ndarray, index: int, count: int, options: dict | None = None) -> np.ndarray:
    """Run one band through iw_process_image (host buffers).  Returns (rows_in_band, bpr) uint8."""
    from . import configs
    opts = dict(options or {})
    opts["cuda:band"] = f"{index}/{count}"
    _, px = configs.run_host(lib, cfg, rows, options=opts)
    return px


def gather_bands(local: np.ndarray, out_h: int, rank: int, world: int, device=None):
    """Gather every rank's band to rank 0 (torch.distributed must be initialised).  Bands may differ in height
    by one row, so they are padded to the tallest band for the collective and trimmed afterwards.
    Returns the assembled (out_h, bpr) array on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    bands = all_bands(out_h, world)
    hmax = max(b - a for a, b in bands)
    bpr = local.shape[1]
    t = torch.zeros((hmax, bpr), dtype=torch.uint8, device=device or "cpu")
    t[: local.shape[0]] = torch.from_numpy(local).to(t.device)
    parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gather_list=parts, dst=0)
    if rank != 0:
        return None
    return np.concatenate([parts[i][: b - a].cpu().numpy() for i, (a, b) in enumerate(bands)], axis=0)
