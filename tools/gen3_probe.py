"""Run one configuration through the job layer on the fused kernel, launch by launch, with a synchronise after every
launch; on an error print the fused kernel's own diagnostics.  python tools/gen3_probe.py [config] [launches] [opt=val ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from imageworsener_b200 import configs, libimagew as L  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
opts = dict(a.split("=", 1) for a in sys.argv[3:])
lib = L.load_product()
dll = lib.dll
dll.iwb200_job_create.restype = C.c_void_p
dll.iwb200_job_create.argtypes = [C.c_void_p, C.c_int]
dll.iwb200_job_destroy.argtypes = [C.c_void_p]
dll.iwb200_job_run_device.restype = C.c_int
dll.iwb200_job_run_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
dll.iwb200_job_last_kernel_generation.argtypes = [C.c_void_p]
dll.iwb200_debug_fused_trap.argtypes = [C.c_char_p, C.c_int]
cfg = configs.get(name)
rows = configs.synth_rows(lib, cfg)
ctx = lib.iw_create_context(None)
img = L.iw_image()
img.imgtype, img.bit_depth, img.sampletype = cfg.imgtype, cfg.bit_depth, 0
img.width, img.height, img.bpr, img.pixels = cfg.in_w, cfg.in_h, cfg.in_bpr, None
lib.iw_set_input_image(ctx, C.byref(img))
for k, v in opts.items():
    lib.iw_set_option(ctx, k.encode(), v.encode())
cfg.apply(lib, ctx)
job = dll.iwb200_job_create(ctx, 0)
assert job, lib.errmsg(ctx)
d_in = torch.from_numpy(rows.reshape(-1)).cuda()
out_bpr = cfg.out_w * L.imgtype_num_channels(cfg.imgtype)
d_out = torch.zeros(cfg.out_h * out_bpr, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
ref = None
for i in range(n):
    ok = dll.iwb200_job_run_device(job, d_in.data_ptr(), cfg.in_bpr, 0, d_out.data_ptr(), out_bpr, 0, cfg.out_h, stream.cuda_stream)
    try:
        stream.synchronize()
    except Exception as e:  # noqa: BLE001
        buf = C.create_string_buffer(4096)
        dll.iwb200_debug_fused_trap(buf, 4096)
        print("launch", i, "failed:", str(e).splitlines()[0], "|", buf.value.decode())
        sys.exit(1)
    h = configs.fnv1a64(lib, d_out.cpu().numpy())
    print("launch", i, "ok", ok, "generation", dll.iwb200_job_last_kernel_generation(job), "fnv %016x" % h, flush=True)
