#!/bin/bash
# Build kernel variants ON THE GPU BOX (quick mode: only the cfg2 instantiations) and time each.
cd "$(dirname "$0")/.."
SRC="imageworsener_b200/csrc"
build() {
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false \
    -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -cudart static -shared -Xlinker -Bsymbolic \
    -DIWB_BUILDING=1 -DIWB_QUICK=1 "$@" -o imageworsener_b200/libimagew_b200.so \
    $SRC/api.cpp $SRC/plan.cpp $SRC/weights.cpp $SRC/tables.cpp $SRC/process.cpp $SRC/pipeline.cu 2>&1 | grep -E "error" 
}
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; print('   step %.3f ms  V %.3f  H %.3f' % (d['ms_per_step'], k['fused_vh_tile']['ms_per_step'], k['h_pass_epilogue']['ms_per_step']))"
}
while read -r name flags; do
  [ -z "$name" ] && continue
  echo "== $name: $flags"
  build $flags; run
done <<'VARIANTS'
base
tb4 -DIWB_V_TB=4
vpipe -DIWB_V_PIPE=1
hpipe -DIWB_H_PIPE=1
rep8 -DIWB_LUT_REP=8
ring4 -DIWB_RING_D=4
ring16 -DIWB_RING_D=16
hw2 -DIWB_H_WARPS=2
hw8 -DIWB_H_WARPS=8
vmin5 -DIWB_V_MINBLOCKS=5
vmin3 -DIWB_V_MINBLOCKS=3
VARIANTS
