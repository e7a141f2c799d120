#!/usr/bin/env python
"""bench.py -- output Mpix/s of the pixel pipeline on BASELINE.json's headline configuration (cfg2:
7680x4320 RGB8 -> 3840x2160, Lanczos3 in linear light, ordered dither, 8-bit sRGB out).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl ours|reference]

A step = one pass of the whole pipeline over one synthetic image.  `value` times the step with the image already
resident in HBM (CUDA events on the launching stream, rotating input/output buffers larger than L2 between
steps); `e2e` times the public C-ABI call iw_process_image() on HOST buffers, so it contains the H2D copy of the
input and the D2H copy of the result.  With N > 1 (launched by torchrun, one rank per GPU) every rank processes its
own image (images shard across GPUs with no data-path collective: weak scaling); the time is the max over ranks.
`--impl reference` times the unmodified reference build (oracle/_ref) on the host cores.
`--config` also takes the other BASELINE.json configurations (cfg3, cfg4auto, cfg4box, cfg5) and a few everyday shapes
(frac, catrom3, up2, up4, gray2: imageworsener_b200/configs.py); the driver's line is always cfg2.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from imageworsener_b200 import configs, libimagew as L  # noqa: E402  (the reference arm uses only the config
#                                                                  # descriptions and the ctypes binding class from it)

METRIC = "output Mpix/s, 8K->4K Lanczos3 linear-light RGB8 + ordered dither"
KERNEL_CLASSES = ["v_pass_generic", "h_pass_epilogue", "prng_fill", "errdiff", "post", "v_pass_stream", "vh_fused"]
# measured FP64 issue rate for independent DMUL+DADD pairs.  MEASURED_PEAKS.json has no FP64 entry: this is the
# builder's own microbenchmark (tools/fp64_peak.cu, profiles/r1_fp64_peak.txt), consistent with 148 SMs x 64 lanes x
# 1.965 GHz = 18.6 T; the JSON line says so (`fp64.peak_source`)
FP64_PEAK_LANE_INSTR = 18.4e12
FP64_PEAK_SOURCE = "builder-measured (tools/fp64_peak.cu, profiles/r1_fp64_peak.txt); not in MEASURED_PEAKS.json"


def kernel_sources_sha() -> str:
    """Fingerprint of the device code of the headline configuration's kernels (the generation-2 streaming kernels:
    stream2*.cu/.cuh/.inc and what they include).  The ncu capture in profiles/traffic.json is only valid for the kernels
    it was taken on."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "imageworsener_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith("stream2") or f in ("kernels.cuh", "prng.cuh"):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel_class: str):
    """DRAM bytes (read + write) per launch of the kernel class from the committed `ncu --set full` capture of this
    same command (profiles/traffic.json, written by tools/ncu_summary.py).  The capture is stamped with the
    fingerprint of the kernel sources it was taken on; None (and why) when HEAD's kernels differ or nothing is recorded."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
    except Exception:  # noqa: BLE001
        return None, "no capture recorded"
    if d.get("kernel_sources_sha") != kernel_sources_sha():
        return None, f"capture is of other kernel sources ({d.get('kernel_sources_sha')}, commit {d.get('commit')})"
    return d.get(kernel_class), f"ncu --set full capture, commit {d.get('commit')}"


def ncu_issue_cycles(kernel_classes, n_sm=148):
    """Issue-port cycles per scheduler that the captured launches of these kernel classes need: an FP64 warp instruction
    holds a scheduler's issue port for two cycles, any other instruction for one (measured: tools/issue_probe.cu,
    profiles/r2h_issue_probe.txt -- a DMUL+DADD pair plus NX independent FADDs costs 4 + NX cycles, not max(4, 2 + NX)).
    Instruction counts come from the same stamped capture as `traffic`; None when HEAD's kernels differ."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
    except Exception:  # noqa: BLE001
        return None
    if d.get("kernel_sources_sha") != kernel_sources_sha():
        return None
    total = 0.0
    for k in kernel_classes:
        n, f = d.get(k + "_warp_instr"), d.get(k + "_warp_instr_fp64")
        if n is None or f is None:
            return None
        total += (n + f) / (4.0 * n_sm)
    return total


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while a timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(gpu_index: int):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE any page-locked buffer is allocated
    (first touch puts the pages there): with several ranks per box the host-to-device copies of the end-to-end
    leg otherwise cross the socket interconnect.  Best effort; returns a short description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id
        dom = torch.cuda.get_device_properties(gpu_index).pci_domain_id
        dev = torch.cuda.get_device_properties(gpu_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "no NUMA information"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"node {node} ({len(cpus)} cpus)"
    except Exception as e:  # noqa: BLE001
        return f"unbound ({type(e).__name__})"
    return "unbound"


def make_ctx(lib: L.Lib, cfg, rows=None, options=None):
    ctx = lib.iw_create_context(None)
    if cfg.max_malloc:
        lib.iw_set_max_malloc(ctx, cfg.max_malloc)
    if rows is not None:
        lib.set_input_from_numpy(ctx, rows, cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
    else:  # header only: pixels live in HBM
        img = L.iw_image()
        img.imgtype, img.bit_depth, img.sampletype = cfg.imgtype, cfg.bit_depth, 0
        img.width, img.height, img.bpr, img.pixels = cfg.in_w, cfg.in_h, cfg.in_bpr, None
        lib.iw_set_input_image(ctx, C.byref(img))
    for k, v in (options or {}).items():
        lib.iw_set_option(ctx, k.encode(), str(v).encode())
    cfg.apply(lib, ctx)
    return ctx


def reference_lib():
    import oracle                        # test infrastructure: the unmodified reference build (oracle/_ref)
    return oracle.load_reference()


def synth_rows_host(cfg, seed=configs.XORSHIFT_SEED):
    """The synthetic input without the product library (oracle/libiworacle.so's plain-C generator)."""
    import oracle
    return oracle.synth_xorshift(cfg.in_bytes, seed).reshape(cfg.in_h, cfg.in_bpr)


def cpu_reference_once(cfg, rows, threads: int):
    """`threads` reference contexts in parallel, one image each (the reference has no intra-image threading).
    Returns seconds."""
    ref = reference_lib()
    errs = []

    def work():
        try:
            configs.run_host(ref, cfg, rows)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    if errs:
        raise errs[0]
    return dt


def workload_config(name, cfg):
    """The `config` keys both arms print (the driver compares them)."""
    return {"workload": f"{name}: {cfg.description}", "in": [cfg.in_w, cfg.in_h], "out": [cfg.out_w, cfg.out_h]}


def run_reference_arm(args, rank, world):
    """The unmodified reference build on the host cores, on the SAME configuration as the GPU arm (full size).  A step
    = `threads` contexts in parallel, one whole image each (the reference cannot thread one image); nothing of the
    product is loaded."""
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = max(1, min(cores, 16))
    cfg = configs.get(args.config)
    rows = synth_rows_host(cfg)
    # bound the run: one 8K image is ~3.5 s on one core, so steps x that must stay within a few minutes
    t_probe = cpu_reference_once(cfg, rows, threads) if args.warmup > 0 else None
    budget_s = 420.0
    steps = args.steps
    if t_probe is not None and t_probe * (args.steps + max(0, args.warmup - 1)) > budget_s:
        steps = max(1, int(budget_s / t_probe) - max(0, args.warmup - 1))
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_once(cfg, rows, threads)
    t = 0.0
    for _ in range(steps):
        t += cpu_reference_once(cfg, rows, threads)
    value = threads * steps * cfg.out_mpix / t
    sample = (f"{threads} reference contexts in parallel (one thread each; the reference has no intra-image "
              f"threading), each resizing one full {cfg.in_w}x{cfg.in_h} image per step; {steps} timed steps"
              + ("" if steps == args.steps else f" (of {args.steps} asked: bounded to ~{budget_s:.0f} s)"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.config, cfg),
            "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)



class JobInfo(C.Structure):
    _fields_ = [("out_width", C.c_int), ("out_height", C.c_int), ("out_imgtype", C.c_int),
                ("out_bit_depth", C.c_int), ("out_sampletype", C.c_int), ("out_channels", C.c_int),
                ("out_bpr", C.c_size_t), ("in_width_phys", C.c_int), ("in_height_phys", C.c_int),
                ("in_bpr", C.c_size_t), ("bandable", C.c_int), ("uses_fused_kernel", C.c_int),
                ("taps_v", C.c_double), ("taps_h", C.c_double)]


def bind_job_api(dll):
    dll.iwb200_kernel_launch_count.restype = C.c_ulonglong
    dll.iwb200_job_create.restype = C.c_void_p
    dll.iwb200_job_create.argtypes = [C.c_void_p, C.c_int]
    dll.iwb200_job_destroy.argtypes = [C.c_void_p]
    dll.iwb200_job_run_device.restype = C.c_int
    dll.iwb200_job_run_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                          C.c_int, C.c_int, C.c_void_p]
    dll.iwb200_job_band_input_rows.restype = C.c_int
    dll.iwb200_job_band_input_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    dll.iwb200_job_set_profiling.argtypes = [C.c_void_p, C.c_int]
    dll.iwb200_job_reset_stats.argtypes = [C.c_void_p]
    dll.iwb200_job_kernel_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    dll.iwb200_job_last_kernel_generation.argtypes = [C.c_void_p]
    dll.iwb200_job_last_error.restype = C.c_char_p
    dll.iwb200_job_last_error.argtypes = [C.c_void_p]
    dll.iwb200_job_get_info.argtypes = [C.c_void_p, C.POINTER(JobInfo)]


def known_answers():
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))
    except Exception:  # noqa: BLE001
        return {}


def measure_config_brief(lib, name, device, steps=5, warmup=3, e2e_steps=2, options=None, zoneplate=False):
    """One other BASELINE configuration on one GPU: device-resident ms per image (CUDA events on the launching stream,
    L2 defeated by buffer rotation or by the image size itself), end to end through iw_process_image on host buffers,
    the whole-step roofline fraction, and the FNV-1a-64 of the output against the reference's recorded fingerprint."""
    import torch
    dll = lib.dll
    cfg = configs.get(name)
    rows = configs.zoneplate_rows(cfg) if zoneplate else configs.synth_rows(lib, cfg)
    opts = {"cuda:device": device, **(options or {})}
    ctx = make_ctx(lib, cfg, rows=None, options=opts)
    job = dll.iwb200_job_create(ctx, device)
    if not job:
        raise SystemExit(f"{name}: job creation failed: " + lib.errmsg(ctx))
    info = JobInfo()
    dll.iwb200_job_get_info(job, C.byref(info))
    out_bytes = info.out_bpr * info.out_height
    per_set = cfg.in_bytes + out_bytes
    nbuf = 1 if per_set > 300e6 else min(8, max(2, int(np.ceil(300e6 / per_set)) + 1))
    host = torch.from_numpy(rows.reshape(-1))
    d_in = [host.to("cuda") for _ in range(nbuf)]
    d_out = [torch.empty(out_bytes, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    stream = torch.cuda.Stream()

    def step(i):
        ok = dll.iwb200_job_run_device(job, d_in[i % nbuf].data_ptr(), cfg.in_bpr, 0, d_out[i % nbuf].data_ptr(), info.out_bpr,
                                       0, info.out_height, stream.cuda_stream)
        if not ok:
            raise SystemExit(f"{name}: run_device failed: " + dll.iwb200_job_last_error(job).decode())

    with torch.cuda.stream(stream):
        for i in range(warmup):
            step(i)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step(i)
        e1.record(stream)
    stream.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    gen = int(dll.iwb200_job_last_kernel_generation(job))
    fnv = "%016x" % configs.fnv1a64(lib, d_out[(steps - 1) % nbuf].cpu().numpy())
    dll.iwb200_job_destroy(job)
    lib.iw_destroy_context(ctx)
    del d_in, d_out
    torch.cuda.empty_cache()
    if zoneplate:      # device time on an image-like input only (no recorded fingerprint, same bytes over PCIe)
        return {"device_ms": dev_ms, "device_mpix_s": cfg.out_mpix / (dev_ms / 1000.0), "input": "1024 x 1024 zone plate per channel, tiled"}

    def e2e_once():
        c = make_ctx(lib, cfg, rows=rows, options=opts)
        t0 = time.perf_counter()
        ok = lib.iw_process_image(c)
        dt = time.perf_counter() - t0
        if not ok:
            raise SystemExit(f"{name}: iw_process_image failed: " + lib.errmsg(c))
        lib.iw_destroy_context(c)
        return dt

    e2e_once()
    e2e_ms = 1000.0 * sum(e2e_once() for _ in range(e2e_steps)) / e2e_steps
    peak, _ = measured_peak_hbm()
    want = known_answers().get(name, {}).get("fnv1a64")
    return {"workload": f"{name}: {cfg.description}", "in": [cfg.in_w, cfg.in_h], "out": [cfg.out_w, cfg.out_h],
            "device_ms": dev_ms, "device_mpix_s": cfg.out_mpix / (dev_ms / 1000.0), "e2e_ms": e2e_ms,
            "e2e_mpix_s": cfg.out_mpix / (e2e_ms / 1000.0), "algorithmic_bytes": per_set,
            "roofline_frac_whole_step": per_set / (dev_ms * 1e-3) / 1e9 / peak, "kernel_generation": gen,
            "fnv1a64": fnv, "fnv_matches_reference": (fnv == want) if want else None,
            "l2_defeat": f"{nbuf} rotating buffer set(s), {nbuf * per_set / 1e6:.0f} MB"}


def measure_bands(lib, name, rank, local_rank, world, dist, steps=5, warmup=3):
    """One large image over all ranks' GPUs: (device) every rank resamples its output-row band from its own
    slab of input rows (+ filter halo), the bands are gathered to rank 0 with NCCL inside the timed step; (end to end)
    rank 0 alone calls iw_process_image with cuda:devices = every GPU of the job -- the library itself splits the image,
    copies each band's rows to its device and the results straight back into the caller's buffer."""
    import torch
    from imageworsener_b200 import bands as _bands
    dll = lib.dll
    cfg = configs.get(name)
    rows = configs.synth_rows(lib, cfg)               # every rank generates the same image, uploads only its slab
    ctx = make_ctx(lib, cfg, rows=None, options={"cuda:device": local_rank})
    job = dll.iwb200_job_create(ctx, local_rank)
    if not job:
        raise SystemExit(f"{name}: job creation failed: " + lib.errmsg(ctx))
    info = JobInfo()
    dll.iwb200_job_get_info(job, C.byref(info))
    r0, r1 = _bands.band_rows(info.out_height, rank, world)
    hmax = max(b - a for a, b in _bands.all_bands(info.out_height, world))
    first, last = C.c_int(0), C.c_int(0)
    dll.iwb200_job_band_input_rows(job, r0, r1, C.byref(first), C.byref(last))
    slab = torch.from_numpy(rows[first.value:last.value + 1].reshape(-1)).to("cuda")
    d_out = torch.empty(hmax * info.out_bpr, dtype=torch.uint8, device="cuda")
    gathered = [torch.empty(hmax * info.out_bpr, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
    stream = torch.cuda.Stream()

    def step():
        ok = dll.iwb200_job_run_device(job, slab.data_ptr(), cfg.in_bpr, first.value, d_out.data_ptr(), info.out_bpr, r0, r1,
                                       stream.cuda_stream)
        if not ok:
            raise SystemExit(f"{name}: run_device failed: " + dll.iwb200_job_last_error(job).decode())
        dist.gather(d_out, gather_list=gathered, dst=0)

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    fnv = None
    if rank == 0:
        parts = []
        for k, (a, b) in enumerate(_bands.all_bands(info.out_height, world)):
            parts.append(gathered[k][: (b - a) * info.out_bpr].cpu().numpy())
        fnv = "%016x" % configs.fnv1a64(lib, np.concatenate(parts))
    dll.iwb200_job_destroy(job)
    lib.iw_destroy_context(ctx)
    del slab, d_out, gathered
    torch.cuda.empty_cache()
    # (the other ranks wait on the CPU: an NCCL barrier would keep a kernel spinning on the very GPUs rank 0 is about to use)
    cpu_group = dist.new_group(backend="gloo")
    dist.barrier(group=cpu_group)

    # end to end: one process, one call, N devices
    e2e_ms = e2e_1 = None
    fnv_e2e = None
    if rank == 0:
        def once(devs):
            c = make_ctx(lib, cfg, rows=rows, options={"cuda:devices": ",".join(str(d) for d in devs)})
            t0 = time.perf_counter()
            ok = lib.iw_process_image(c)
            dt = time.perf_counter() - t0
            if not ok:
                raise SystemExit(f"{name}: iw_process_image (cuda:devices) failed: " + lib.errmsg(c))
            img, px = lib.get_output_numpy(c)
            lib.iw_destroy_context(c)
            return dt, px
        devs = list(range(world))
        once(devs)
        ts = []
        for _ in range(3):
            dt, px = once(devs)
            ts.append(dt)
        e2e_ms = 1000.0 * sum(ts) / len(ts)
        fnv_e2e = "%016x" % configs.fnv1a64(lib, px)
        once([0])
        e2e_1 = 1000.0 * once([0])[0]
    dist.barrier(group=cpu_group)
    if rank != 0:
        return None
    want = known_answers().get(name, {}).get("fnv1a64")
    return {"workload": f"{name}: {cfg.description}", "in": [cfg.in_w, cfg.in_h], "out": [cfg.out_w, cfg.out_h], "bands": world,
            "device_ms": dev_ms, "device_mpix_s": cfg.out_mpix / (dev_ms / 1000.0),
            "device_note": "per step: every rank's band kernels + NCCL gather of the bands to rank 0, max over ranks",
            "e2e_ms": e2e_ms, "e2e_mpix_s": cfg.out_mpix / (e2e_ms / 1000.0), "e2e_ms_one_device": e2e_1,
            "e2e_note": "iw_process_image on rank 0 with cuda:devices=0..N-1 (host buffers; bands copied to / from N devices in parallel)",
            "fnv1a64_device": fnv, "fnv1a64_e2e": fnv_e2e,
            "fnv_matches_reference": (fnv == want and fnv_e2e == want) if want else None}


def measure_batch(lib, name, rank, local_rank, world, dist, total=1024, distinct=64, depth=4):
    """BASELINE configuration 4: `total` same-shaped images (`distinct` different noise images, cycled) through the batch
    entry point (iwb200_batch_*), one image per GPU round-robin over the ranks: page-locked host buffers, the copies and
    kernels of consecutive images overlapped inside the library.  Every distinct image's result is compared with the
    fingerprint of the first one's known answer where recorded, and all copies of an image must agree."""
    import torch
    dll = lib.dll
    dll.iwb200_batch_create.restype = C.c_void_p
    dll.iwb200_batch_create.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int]
    dll.iwb200_batch_submit.restype = C.c_int
    dll.iwb200_batch_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    dll.iwb200_batch_wait.restype = C.c_int
    dll.iwb200_batch_wait.argtypes = [C.c_void_p]
    dll.iwb200_batch_destroy.argtypes = [C.c_void_p]
    dll.iwb200_batch_last_error.restype = C.c_char_p
    dll.iwb200_batch_last_error.argtypes = [C.c_void_p]
    cfg = configs.get(name)
    mine = total // world
    nd = min(distinct, mine)
    ins = torch.empty((nd, cfg.in_bytes), dtype=torch.uint8).pin_memory()
    for k in range(nd):
        ins[k].copy_(torch.from_numpy(configs.synth_rows(lib, cfg, seed=configs.XORSHIFT_SEED + 1000 * rank + k).reshape(-1)))
    ctx = make_ctx(lib, cfg, rows=None, options={"cuda:device": local_rank})
    devs = (C.c_int * 1)(local_rank)
    b = dll.iwb200_batch_create(ctx, devs, 1, depth)
    if not b:
        raise SystemExit(f"{name}: batch creation failed: " + lib.errmsg(ctx))
    out_bytes = cfg.out_w * cfg.out_h * cfg.nch
    outs = torch.zeros((mine, out_bytes), dtype=torch.uint8).pin_memory()

    def run(n):
        for i in range(n):
            if not dll.iwb200_batch_submit(b, ins[i % nd].data_ptr(), cfg.in_bpr, outs[i].data_ptr(), cfg.out_w * cfg.nch):
                raise SystemExit(f"{name}: batch submit failed: " + dll.iwb200_batch_last_error(b).decode())
        if not dll.iwb200_batch_wait(b):
            raise SystemExit(f"{name}: batch wait failed: " + dll.iwb200_batch_last_error(b).decode())

    run(min(mine, 64))                                   # warm (plan, pools, clocks)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(mine)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    # every copy of a distinct image gives the same bytes; the first image's fingerprint is the recorded known answer
    o = outs.numpy()
    consistent = all(np.array_equal(o[i], o[i % nd]) for i in range(mine))
    fnv0 = "%016x" % configs.fnv1a64(lib, o[0])
    want = known_answers().get(name, {}).get("fnv1a64") if rank == 0 else None
    dll.iwb200_batch_destroy(b)
    lib.iw_destroy_context(ctx)
    per_img = cfg.in_bytes + out_bytes
    return {"workload": f"{name}: {cfg.description}", "images": mine * world, "distinct_images_per_gpu": nd, "slots_per_gpu": depth,
            "seconds": dt, "images_per_s": mine * world / dt, "mpix_s": mine * world * cfg.out_mpix / dt,
            "pcie_bytes_per_image": per_img, "host_gb_s": mine * world * per_img / dt / 1e9,
            "copies_of_an_image_agree": bool(consistent),
            "first_image_fnv_matches_reference": (fnv0 == want) if want else None,
            "api": "iwb200_batch_submit (page-locked host buffers; H2D + kernels + D2H of consecutive images overlapped)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="context option name=value (tuning)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the other BASELINE configurations / the band block")
    ap.add_argument("--shard", default="images", choices=["images", "bands"],
                    help="N>1: one image per GPU (weak scaling, default) or one image split into output-row bands "
                         "gathered to rank 0 over NCCL (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib = L.load_product()
    dll = lib.dll
    bind_job_api(dll)
    Info = JobInfo

    cfg = configs.get(args.config)
    rows = configs.synth_rows(lib, cfg, seed=configs.XORSHIFT_SEED + rank)

    # ---------------- device-resident throughput -------------------------------------------------------
    extra = dict(o.split("=", 1) for o in args.opt)
    ctx = make_ctx(lib, cfg, rows=None, options={"cuda:device": local_rank, **extra})
    job = dll.iwb200_job_create(ctx, local_rank)
    if not job:
        raise SystemExit("job creation failed: " + lib.errmsg(ctx))
    info = Info()
    dll.iwb200_job_get_info(job, C.byref(info))
    out_bytes = info.out_bpr * info.out_height
    nbuf = max(2, int(np.ceil(300e6 / (cfg.in_bytes + out_bytes))) + 1)   # rotate through > 2x L2 of data
    nbuf = min(nbuf, 8)
    host = torch.from_numpy(rows.reshape(-1))
    d_in = []
    for b in range(nbuf):
        t = host.to("cuda", non_blocking=False)
        if b:
            t = torch.roll(t, shifts=b * 4099)       # different bytes per buffer, same statistics
        d_in.append(t.contiguous())
    d_out = [torch.empty(out_bytes, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    stream = torch.cuda.Stream()

    from imageworsener_b200 import bands as _bands
    band_mode = args.shard == "bands" and world > 1
    r0, r1 = _bands.band_rows(info.out_height, rank, world) if band_mode else (0, info.out_height)
    hmax = max(b - a for a, b in _bands.all_bands(info.out_height, world)) if band_mode else info.out_height
    gathered = [torch.empty(hmax * info.out_bpr, dtype=torch.uint8, device="cuda") for _ in range(world)] \
        if band_mode and rank == 0 else None

    def step(i):
        out_t = d_out[i % nbuf]
        ok = dll.iwb200_job_run_device(job, d_in[i % nbuf].data_ptr(), cfg.in_bpr, 0, out_t.data_ptr(),
                                       info.out_bpr, r0, r1, stream.cuda_stream)
        if not ok:
            raise SystemExit("run_device failed: " + dll.iwb200_job_last_error(job).decode())
        if band_mode:   # the only exchange of the path: output bands -> rank 0 (NCCL over NVLink)
            dist.gather(out_t[: hmax * info.out_bpr], gather_list=gathered, dst=0)

    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = dll.iwb200_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = dll.iwb200_kernel_launch_count() - l0
    # keep the sampler running a little longer if the region was very short, so it sees load
    if ms_total < 1500:
        # (a step count every rank agrees on: in band mode each step contains a collective)
        extra_steps = int(min(20000, max(8, 1500.0 / max(ms_total / args.steps, 1e-3))))
        with torch.cuda.stream(stream):
            for i in range(extra_steps):
                step(i)
                if i % 8 == 7:
                    stream.synchronize()
        stream.synchronize()
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    value = (1 if band_mode else world) * cfg.out_mpix / (ms_per_step / 1000.0)

    # ---------------- per-kernel device time -> roofline of the dominant kernel ------------------------
    dll.iwb200_job_set_profiling(job, 1)
    dll.iwb200_job_reset_stats(job)
    prof_steps = min(5, args.steps)
    with torch.cuda.stream(stream):
        for i in range(prof_steps):
            step(i)
    stream.synchronize()
    stats = {}
    for k, nm in enumerate(KERNEL_CLASSES):
        n, ms = C.c_int(0), C.c_double(0.0)
        dll.iwb200_job_kernel_stats(job, k, C.byref(n), C.byref(ms))
        if n.value:
            stats[nm] = {"launches_per_step": n.value / prof_steps, "ms_per_step": ms.value / prof_steps}
    dll.iwb200_job_set_profiling(job, 0)
    kernel_generation = int(dll.iwb200_job_last_kernel_generation(job))
    dom = max(stats, key=lambda k: stats[k]["ms_per_step"])
    alg_bytes_step = cfg.in_bytes + out_bytes                        # read input once + write output once
    lp = stats[dom]["launches_per_step"]
    avg_launch_ms = stats[dom]["ms_per_step"] / lp
    achieved = (alg_bytes_step / lp) / (avg_launch_ms * 1e-3) / 1e9   # GB/s, algorithmic bytes of the rows a launch covers
    peak, peak_src = measured_peak_hbm()
    taps = info.taps_v + info.taps_h
    step_gbs = alg_bytes_step / (ms_per_step * 1e-3) / 1e9
    fp64_rate = 2.0 * taps / (ms_per_step * 1e-3)
    traffic, traffic_src = ncu_traffic(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_step": alg_bytes_step, "launches_per_step": lp,
                "avg_launch_ms": avg_launch_ms, "kernel_ms_per_step": stats,
                "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak},
                "fp64_tap_pairs_per_step": taps,
                "fp64": {"achieved_lane_instr_per_s": fp64_rate, "peak_lane_instr_per_s": FP64_PEAK_LANE_INSTR,
                         "peak_source": FP64_PEAK_SOURCE, "frac": fp64_rate / FP64_PEAK_LANE_INSTR,
                         "hbm_frac_ceiling": (alg_bytes_step / peak / 1e9) / (2.0 * taps / FP64_PEAK_LANE_INSTR)},
                "note": "every tap is one DMUL + one DADD (FMA forbidden by bit-exactness): the FP64 pipe binds before "
                        "HBM, `fp64.hbm_frac_ceiling` is the HBM fraction at 100 % FP64 issue; see DESIGN.md"}
    issue_cycles = ncu_issue_cycles(list(stats)) if args.config == "cfg2" else None
    if issue_cycles:
        # the tighter roof for these kernels: issue-port cycles (2 per FP64 instruction, 1 per other) of the step's
        # captured instruction streams against the step's cycles at the SM clock sampled during the timed region
        mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        roofline["issue_port"] = {"cycles_needed_per_scheduler": issue_cycles, "cycles_elapsed": ms_per_step * 1e-3 * mhz * 1e6,
                                  "frac": issue_cycles / (ms_per_step * 1e-3 * mhz * 1e6),
                                  "model": "2 issue cycles per FP64 warp instruction + 1 per other (tools/issue_probe.cu); "
                                           "instruction counts from the stamped ncu capture"}
    dll.iwb200_job_destroy(job)
    lib.iw_destroy_context(ctx)

    # ---------------- end to end through iw_process_image on host buffers ------------------------------
    def e2e_once():
        c = make_ctx(lib, cfg, rows=rows, options={"cuda:device": local_rank, **extra})
        t0 = time.perf_counter()
        ok = lib.iw_process_image(c)
        dt = time.perf_counter() - t0
        if not ok:
            raise SystemExit("iw_process_image failed: " + lib.errmsg(c))
        img = L.iw_image()
        lib.iw_get_output_image(c, C.byref(img))
        first = C.cast(img.pixels, C.POINTER(C.c_uint8))[0]     # the result is host-resident
        lib.iw_destroy_context(c)
        return dt, first

    for _ in range(2):
        e2e_once()
    barrier()
    e2e_t = 0.0
    for _ in range(args.e2e_steps):
        dt, _ = e2e_once()
        e2e_t += dt
    barrier()
    e2e_ms = max_over_ranks(1000.0 * e2e_t / args.e2e_steps)
    e2e = {"value": world * cfg.out_mpix / (e2e_ms / 1000.0), "unit": "Mpix/s", "h2d_bytes_per_step": cfg.in_bytes,
           "d2h_bytes_per_step": out_bytes, "ms_per_step": e2e_ms, "steps": args.e2e_steps,
           "api": "iw_process_image (host buffers; plan + H2D + kernels + D2H inside the timed call)"}

    # ---------------- CPU baseline: the reference build on this box's host cores ------------------------
    cpu = None
    import oracle                                                    # (the checker / CPU baseline; never timed as ours)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(oracle.REFERENCE_SO):
        dt = cpu_reference_once(cfg, rows, 1)
        cpu = {"value": cfg.out_mpix / dt, "unit": "Mpix/s", "cores": 1, "kind": "reference",
               "sample": f"one full {cfg.in_w}x{cfg.in_h} image through the unmodified reference build, 1 thread "
                         f"(it has no intra-image threading), {dt:.1f} s", "host_cores_available": os.cpu_count()}

    # ---------------- the other BASELINE configurations (N = 1) / one large image in bands over all GPUs (N > 1) ------
    other, bands_block, batch_block = None, None, None
    if not args.no_other_configs and args.config == "cfg2":
        if world == 1:
            other = {}
            for nm in ("cfg3", "cfg4auto", "cfg4box", "cfg5"):
                try:
                    other[nm] = measure_config_brief(lib, nm, local_rank)
                except SystemExit as e:  # noqa: PERF203
                    other[nm] = {"error": str(e)}
            # the same kernels on an image-like input (noise is the worst case for every table lookup and
            # clips after filters with negative lobes); device time only
            for nm in ("cfg2", "cfg3", "cfg5"):
                try:
                    other[nm + "_zoneplate"] = measure_config_brief(lib, nm, local_rank, zoneplate=True)
                except SystemExit as e:  # noqa: PERF203
                    other[nm + "_zoneplate"] = {"error": str(e)}
        else:
            bands_block = measure_bands(lib, "cfg5", rank, local_rank, world, dist)
        batch_block = {nm: measure_batch(lib, nm, rank, local_rank, world, dist if world > 1 else None) for nm in ("cfg4auto", "cfg4box")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if band_mode else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {**workload_config(args.config, cfg), "images_per_step_per_gpu": 1,
                           "sharding": ("one image split into output-row bands, gathered to rank 0 with NCCL" if band_mode else
                                        "one image per GPU, no data-path collective") if world > 1 else "single GPU",
                           "l2_defeat": f"{nbuf} rotating input/output buffer sets ({nbuf * (cfg.in_bytes + out_bytes) / 1e6:.0f} MB > 126 MB L2)",
                           "kernel_generation": kernel_generation, "host_numa_binding": numa},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu}
        if other is not None:
            line["other_configs"] = other
        if bands_block is not None:
            line["bands"] = bands_block
        if batch_block is not None:
            line["batch"] = batch_block
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
