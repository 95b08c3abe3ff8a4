#!/bin/bash
# per-kernel step profile of a debug-build variant with a SMB_WS_DBG mask:  tools/gpu_dbgprof.sh <variant> "<masks>"
mkdir -p gpurun_out
cp _variants/$1.so shapemol_b200/libshapemol_b200.so
for m in $2; do
  echo "== variant $1 SMB_WS_DBG=$m"
  SMB_WS_DBG=$m timeout 300 python tools/prof_step.py --mols 16384 --fixed-atoms 27 2>&1 | grep "step\|edge_" | tee gpurun_out/dbgprof_$1_$m.txt
done
