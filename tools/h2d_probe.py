"""What the box can do: concurrent host-to-device (and device-to-host) copies from page-locked memory on N GPUs at once.

The end-to-end numbers of `bench.py` at N > 1 are bound by this, not by kernels (e2e efficiency 0.52 at
N = 8).  One thread per GPU, each with its own pinned buffer (first-touched by that thread) and stream, all released by
a barrier; reports GB/s per GPU and in total for N = 1, 2, 4, 8 (as many as the box has), for the cfg2 image size
(99.5 MB in, 24.9 MB out) and for 1 GiB.

    python tools/h2d_probe.py [--json out.json]
"""
import json
import sys
import threading
import time

import torch


def run(n, nbytes_in, nbytes_out, reps=6):
    bar = threading.Barrier(n + 1)
    res = [None] * n

    def worker(i):
        torch.cuda.set_device(i)
        h_in = torch.empty(nbytes_in, dtype=torch.uint8).pin_memory()
        h_in.fill_(i + 1)                       # first touch by the thread that uses it
        h_out = torch.empty(max(nbytes_out, 1), dtype=torch.uint8).pin_memory()
        d_in = torch.empty(nbytes_in, dtype=torch.uint8, device=f"cuda:{i}")
        d_out = torch.zeros(max(nbytes_out, 1), dtype=torch.uint8, device=f"cuda:{i}")
        s = torch.cuda.Stream(device=i)
        with torch.cuda.stream(s):
            d_in.copy_(h_in, non_blocking=True)
            h_out.copy_(d_out, non_blocking=True)
        s.synchronize()
        bar.wait()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for _ in range(reps):
                d_in.copy_(h_in, non_blocking=True)
                if nbytes_out:
                    h_out.copy_(d_out, non_blocking=True)
        s.synchronize()
        res[i] = time.perf_counter() - t0
        bar.wait()

    th = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
    for t in th:
        t.start()
    bar.wait()
    bar.wait()
    for t in th:
        t.join()
    per = [reps * (nbytes_in + nbytes_out) / r / 1e9 for r in res]
    return {"n_gpus": n, "bytes_in": nbytes_in, "bytes_out": nbytes_out, "gb_s_per_gpu_min": min(per), "gb_s_per_gpu_max": max(per),
            "gb_s_total": reps * (nbytes_in + nbytes_out) * n / max(res) / 1e9, "ms_per_image_slowest": max(res) / reps * 1e3}


def main():
    ngpu = torch.cuda.device_count()
    out = []
    for n in (1, 2, 4, 8):
        if n > ngpu:
            break
        for label, bi, bo in (("cfg2 image in + out (serial on one stream)", 99532800, 24883200), ("1 GiB in", 1 << 30, 0)):
            r = run(n, bi, bo)
            r["what"] = label
            out.append(r)
            print(f"N={n}  {label:45s} {r['gb_s_per_gpu_min']:6.1f}-{r['gb_s_per_gpu_max']:5.1f} GB/s per GPU, {r['gb_s_total']:7.1f} GB/s total, "
                  f"{r['ms_per_image_slowest']:7.2f} ms per transfer set on the slowest GPU", flush=True)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
