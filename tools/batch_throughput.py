"""Batch throughput through the public API (BASELINE cfg4: many small images): T host threads, one context per image,
each calling iw_process_image() on its own host buffers -- the way SURVEY section 8(e) says the reference gets
parallelism for a batch (distinct contexts run concurrently).

    python tools/batch_throughput.py [cfg4auto] [--images 1024] [--threads 1 2 4 8]
"""
import argparse
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageworsener_b200 import configs, libimagew as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="cfg4auto")
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--threads", type=int, nargs="*", default=[1, 2, 4, 8])
    ap.add_argument("--distinct", type=int, default=16, help="distinct synthetic images cycled through")
    args = ap.parse_args()
    prod = L.load_product()
    cfg = configs.get(args.config)
    images = [configs.synth_rows(prod, cfg, seed=configs.XORSHIFT_SEED + i) for i in range(args.distinct)]

    def one(i):
        ctx = prod.iw_create_context(None)
        prod.set_input_from_numpy(ctx, images[i % len(images)], cfg.imgtype, cfg.bit_depth, cfg.in_w, cfg.in_h)
        cfg.apply(prod, ctx)                 # (Python-side option parsing: outside the clock, it would serialise on the GIL)
        return ctx

    def run(ctx):
        ok = prod.iw_process_image(ctx)
        assert ok, prod.errmsg(ctx)
        prod.iw_destroy_context(ctx)

    for i in range(8):                       # warm-up: CUDA context, plan cache, pools
        run(one(i))
    for nt in args.threads:
        # contexts are set up before the clock starts (decoding into iw_malloc_large memory is the codec's job, not the path's)
        ctxs = [one(i) for i in range(args.images)]
        chunks = [ctxs[t::nt] for t in range(nt)]
        ths = [threading.Thread(target=lambda c=c: [run(x) for x in c]) for c in chunks]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        print(f"{cfg.name}: {args.images} images, {nt} threads: {dt * 1e3:.1f} ms = {dt / args.images * 1e3:.4f} ms/image, "
              f"{args.images / dt:.0f} images/s, {args.images * cfg.out_mpix / dt:.1f} out-Mpix/s, "
              f"{args.images * cfg.in_bytes / dt / 1e9:.1f} GB/s of input", flush=True)


if __name__ == "__main__":
    main()
