#!/bin/bash
# On the GPU box: time every tuning build under imageworsener_b200/_variants (tools/build_variants.sh) on the headline
# configuration and check its bytes (tools/check_gen2.py --quick).  Prints one line per variant.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for so in imageworsener_b200/_variants/*.so; do
  name=$(basename "$so" .so)
  par=$(IWB200_LIB=$PWD/$so python tools/check_gen2.py --quick 2>&1 | grep -c "gen2: 0 differing")
  IWB200_LIB=$PWD/$so python bench.py --no-cpu-baseline --no-other-configs --steps 20 --warmup 3 --e2e-steps 1 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - "$name" "$par" gpurun_out/var_$name.json <<'PY'
import json, sys
name, par, path = sys.argv[1:4]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    k = d["roofline"]["kernel_ms_per_step"]
    print(f"{name:12s} parity_ok={par} step {d['ms_per_step']:.4f} ms  " + "  ".join(f"{a} {b['ms_per_step']:.4f}" for a, b in k.items()), flush=True)
except Exception as e:
    print(name, "FAILED", e, flush=True)
PY
done
