#!/bin/bash
# per-kernel step profile of several library variants (no traces):  tools/gpu_varprof.sh "<names>"
mkdir -p gpurun_out
for v in $1; do
  cp _variants/$v.so shapemol_b200/libshapemol_b200.so
  echo "== variant $v"
  timeout 300 python tools/prof_step.py --mols 16384 --fixed-atoms 27 2>&1 | grep "step\|edge_\|node_" | tee gpurun_out/varprof_${v}_27.txt
  timeout 300 python tools/prof_step.py --mols 5000 --fixed-atoms 0 2>&1 | grep "step\|edge_\|node_" | tee gpurun_out/varprof_${v}_prior.txt
done
