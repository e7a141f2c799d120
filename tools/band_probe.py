"""Device time of ONE output-row band of cfg5 on one GPU (no gather): python tools/band_probe.py [nbands]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imageworsener_b200 import configs, libimagew as L, bands as _bands
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = L.load_product(); dll = lib.dll; bench.bind_job_api(dll)
cfg = configs.get("cfg5"); rows = configs.synth_rows(lib, cfg)
ctx = bench.make_ctx(lib, cfg, rows=None, options={"cuda:device": 0})
job = dll.iwb200_job_create(ctx, 0)
info = bench.JobInfo(); dll.iwb200_job_get_info(job, C.byref(info))
for k in (0, nb // 2):
    r0, r1 = _bands.band_rows(info.out_height, k, nb)
    first, last = C.c_int(0), C.c_int(0)
    dll.iwb200_job_band_input_rows(job, r0, r1, C.byref(first), C.byref(last))
    slab = torch.from_numpy(rows[first.value:last.value + 1].reshape(-1)).to("cuda")
    d_out = torch.empty((r1 - r0) * info.out_bpr, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    def step():
        assert dll.iwb200_job_run_device(job, slab.data_ptr(), cfg.in_bpr, first.value, d_out.data_ptr(), info.out_bpr, r0, r1, st.cuda_stream)
    with torch.cuda.stream(st):
        for _ in range(3): step()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10): step()
        e1.record(st)
    st.synchronize()
    print(f"band {k} of {nb}: rows [{r0}, {r1}) device {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
