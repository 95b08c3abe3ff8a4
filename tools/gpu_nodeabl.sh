#!/bin/bash
# node_tc5 ablations on a debug-build variant: SMB_NODE_DBG bit 1 = no weight streaming, 2 = no global stores, 4 = no activation loads
mkdir -p gpurun_out
cp _variants/$1.so shapemol_b200/libshapemol_b200.so
for m in 0 1 2 4 7; do
  echo "== SMB_NODE_DBG=$m"
  SMB_NODE_DBG=$m timeout 300 python tools/prof_step.py --mols 16384 --fixed-atoms 27 2>&1 | grep "node_" | tee gpurun_out/nodeabl_$m.txt
done
