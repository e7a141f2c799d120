#!/bin/bash
# compute-sanitizer sweep over one reduced job per kernel family (on the GPU box): memcheck (out-of-bounds / misaligned
# accesses), synccheck (barrier misuse) and racecheck (shared-memory hazards).  One summary line per run.
#   tools/sanitize.sh [memcheck synccheck racecheck]
cd "$(dirname "$0")/.."
TOOLS=${@:-memcheck synccheck racecheck}
# config, scale, options: streaming kernels (uniform + general paths, gray as RGB, random dither), 16-bit vertical kernel,
# gather-form kernels (alpha, level model), tile kernels (alpha / no alpha with the scatter-form vertical kernel), fused
# kernel, generation 1, error diffusion, thumbnails, enlargement (12 slots)
JOBS=("cfg2 16" "cfg2 16 cuda:uniform=0" "cfg2 16 cuda:gen=1" "cfg2 16 cuda:gen=3" "gray2 16" "cfg5 64" "cfg3 8" "rgba_half 16"
      "rgb16_half 16" "errdiff 32" "cfg4auto 2" "cfg4box 2" "up2 8" "frac 16" "catrom3 16")
for tool in $TOOLS; do
  for job in "${JOBS[@]}"; do
    out=$(timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python tools/run_small.py $job 2>&1)
    rc=$?
    sum=$(echo "$out" | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)
    diff=$(echo "$out" | grep "differing bytes" | tail -1)
    echo "$tool | $job | rc=$rc | $sum | $diff"
    if [ $rc -ne 0 ]; then echo "$out" | grep -E "=========" | head -20; fi
  done
done
