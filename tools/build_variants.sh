#!/bin/bash
# Build tuning variants of the library (quick mode: only the headline configuration's kernels) into
# imageworsener_b200/_variants/<name>.so; they travel to the GPU box and are selected with IWB200_LIB.
#   tools/build_variants.sh name1 "DEF1=.. DEF2=.." name2 "..." ...
cd "$(dirname "$0")/.."
mkdir -p imageworsener_b200/_variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  ( IWB_QUICK=1 IWB_DEFS="$defs" python -c "
from imageworsener_b200 import _build
print(_build.build_product(force=True, out='imageworsener_b200/_variants/$name.so'))" ) &
done
wait
