#!/bin/bash
# compares library variants staged under _variants/ (debug builds): role trace + per-kernel step profile of each
#   tools/gpu_variants.sh "<names>" "<roles>"
mkdir -p gpurun_out
for v in $1; do
  cp _variants/$v.so shapemol_b200/libshapemol_b200.so
  for role in ${2:-1}; do
    echo "== variant $v role $role"; WS_TRACE_N=27 SMB_WS_DBG=$((16 + role * 256)) timeout 120 python tools/ws_trace.py 2>&1 | tail -24 | tee gpurun_out/var_${v}_trace_role${role}.txt
  done
  timeout 300 python tools/prof_step.py --mols 16384 --fixed-atoms 27 2>&1 | grep "step\|edge_" | tee gpurun_out/var_${v}_prof27.txt
  timeout 300 python tools/prof_step.py --mols 5000 --fixed-atoms 0 2>&1 | grep "step\|edge_" | tee gpurun_out/var_${v}_prof_prior.txt
done
