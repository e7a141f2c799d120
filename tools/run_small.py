"""Run one reduced configuration through the CUDA library (for compute-sanitizer / debugging)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageworsener_b200 import configs, libimagew as L
import oracle  # noqa: E402  (test infrastructure: the reference build / port)
name, scale = sys.argv[1], int(sys.argv[2])
opts = dict(o.split("=") for o in sys.argv[3:])
prod = L.load_product()
cfg = configs.get(name, scale)
rows = configs.synth_rows(prod, cfg)
_, a = configs.run_host(prod, cfg, rows, options=opts)
if os.path.exists(oracle.REFERENCE_SO):
    _, b = configs.run_host(oracle.load_reference(), cfg, rows)
    print("differing bytes:", int(np.count_nonzero(a != b)), "of", a.size)
