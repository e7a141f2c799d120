#!/bin/bash
# Capture the round's profiler evidence on the GPU box: launch list of a bench run + full captures of the two
# fast-path kernels.  Outputs under gpurun_out/ (copied into profiles/ by hand together with the summaries).
cd "$(dirname "$0")/.."
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iwb_vstream -s 3 -c 1 -o gpurun_out/${TAG}_vstream \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_vstream.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iwb_hstream -s 3 -c 1 -o gpurun_out/${TAG}_hstream \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_hstream.log 2>&1
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
ls -la gpurun_out | tail -8
