#!/bin/bash
# Capture the round's profiler evidence on the GPU box: launch list of a bench run, full captures of the kernels that carry
# the BASELINE configurations, bench lines of every configuration.  Outputs under gpurun_out/ (summaries are copied into
# profiles/ with tools/ncu_summary.py on the build box).
cd "$(dirname "$0")/.."
TAG=${1:-r2}
PART=${2:-all}   # a / b / all: gpurun brings back at most 64 MiB per call, a full capture is 12-15 MB
B="python bench.py --no-cpu-baseline --e2e-steps 1 --no-other-configs"
if [ $PART != b ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $B --steps 3 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
fi
if [ $PART != b ]; then ncu --set full --clock-control none --import-source on -k regex:iwb_v2 -s 3 -c 1 -o gpurun_out/${TAG}_v2 $B --steps 2 --warmup 3 > gpurun_out/${TAG}_v2.log 2>&1; fi
if [ $PART != b ]; then ncu --set full --clock-control none --import-source on -k regex:iwb_h2 -s 3 -c 1 -o gpurun_out/${TAG}_h2 $B --steps 2 --warmup 3 > gpurun_out/${TAG}_h2.log 2>&1; fi
if [ $PART != a ]; then ncu --set full --clock-control none --import-source on -k regex:iwb_vh3 -s 3 -c 1 -o gpurun_out/${TAG}_vh3 $B --steps 2 --warmup 3 --opt cuda:gen=3 > gpurun_out/${TAG}_vh3.log 2>&1; fi
if [ $PART != a ]; then ncu --set full --clock-control none --import-source on -k regex:iwg_h -s 2 -c 1 -o gpurun_out/${TAG}_cfg3_h $B --config cfg3 --steps 2 --warmup 3 > gpurun_out/${TAG}_cfg3_h.log 2>&1; fi
if [ $PART != b ]; then ncu --set full --clock-control none --import-source on -k regex:iwb_v2 -s 2 -c 1 -o gpurun_out/${TAG}_cfg5_v $B --config cfg5 --steps 2 --warmup 3 > gpurun_out/${TAG}_cfg5_v.log 2>&1; fi
if [ $PART != b ]; then python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json; fi
if [ $PART != b ]; then python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --opt cuda:gen=3 2>&1 | tail -1 > gpurun_out/${TAG}_bench_gen3.json; fi
if [ $PART != b ]; then python bench.py --config errdiff --steps 3 --warmup 3 --e2e-steps 1 --no-other-configs 2>&1 | tail -1 > gpurun_out/${TAG}_bench_errdiff.json; fi
if [ $PART != b ]; then timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference.json; fi
if [ $PART != b ]; then tools/fp64_peak > gpurun_out/${TAG}_fp64_peak.txt 2>&1; fi
ls -la gpurun_out | tail -12
