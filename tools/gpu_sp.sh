#!/bin/bash
mkdir -p gpurun_out
cp _variants/sp.so shapemol_b200/libshapemol_b200.so
for m in 0 64; do
  echo "== SMB_WS_DBG=$m"
  SMB_WS_DBG=$m timeout 300 python tools/prof_step.py --mols 16384 --fixed-atoms 27 2>&1 | grep "step\|edge_"
done
echo "== role 2 trace"; WS_TRACE_N=27 SMB_WS_DBG=$((16 + 2 * 256)) timeout 120 python tools/ws_trace.py 2>&1 | tail -22
