#!/bin/bash
# same-box A/B of release-build library variants: bf16 parity check of each, then the per-kernel step profile, twice
#   tools/gpu_ab.sh "<variant names under _variants/>"
mkdir -p gpurun_out
for v in $1; do
  cp _variants/$v.so shapemol_b200/libshapemol_b200.so
  echo "== check $v"; timeout 180 python tools/check_edge_impl.py 2>&1 | tail -7
done
bash tools/gpu_varprof.sh "$1 $1"
