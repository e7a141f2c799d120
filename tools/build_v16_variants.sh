#!/bin/bash
# Tuning builds of the 16-bit vertical kernel only: stream2_v16.cu recompiled with the given defines and linked with the
# objects of the last full build into imageworsener_b200/_variants/<name>.so (select with IWB200_LIB on the GPU box).
#   tools/build_v16_variants.sh name1 "-DX=1 -DY=2" name2 "..." ...
cd "$(dirname "$0")/.."
C=imageworsener_b200/csrc; O=imageworsener_b200/_obj; V=imageworsener_b200/_variants
mkdir -p $V
OBJS=$(python -c "from imageworsener_b200 import _build as b; import os; print(' '.join(os.path.join('$O', os.path.splitext(n)[0] + '.o') for n in b.DEVICE_SOURCES + b.HOST_SOURCES if n != 'stream2_v16.cu'))")
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  ( ( [ -f $V/$name.o ] && [ "$REUSE" = 1 ] || nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden \
      -DIWB_BUILDING=1 $defs -c $C/stream2_v16.cu -o $V/$name.o ) && \
    nvcc -gencode arch=compute_100a,code=sm_100a -cudart static -shared -Xlinker -Bsymbolic -Xlinker --version-script=$C/exports.map \
      -o $V/$name.so $V/$name.o $OBJS && echo built $V/$name.so ) &
done
wait
