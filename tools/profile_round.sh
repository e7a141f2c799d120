#!/bin/bash
# Capture the round's profiler evidence on the GPU box: launch list of a bench run + full captures of the two
# fast-path kernels.  Outputs under gpurun_out/ (summaries are copied into profiles/ with tools/ncu_summary.py).
cd "$(dirname "$0")/.."
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iwb_v2 -s 3 -c 1 -o gpurun_out/${TAG}_v2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_v2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iwb_h2 -s 3 -c 1 -o gpurun_out/${TAG}_h2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_h2.log 2>&1
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
tools/fp64_peak > gpurun_out/${TAG}_fp64_peak.txt 2>&1
ls -la gpurun_out | tail -8
