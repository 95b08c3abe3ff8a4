#!/bin/bash
# role timeline of CTA 0 (needs a -DSMB_DEBUG build: SMB_NVCC_EXTRA=-DSMB_DEBUG python -m shapemol_b200.build --force)
#   tools/gpu_trace.sh "<roles>" [extra dbg bits]
mkdir -p gpurun_out
for role in ${1:-1}; do
  echo "== role $role dbg ${2:-0}, 27-atom molecules"; WS_TRACE_N=27 SMB_WS_DBG=$((16 + ${2:-0} + role * 256)) timeout 120 python tools/ws_trace.py 2>&1 | tail -${3:-40} | tee gpurun_out/ws_trace_role${role}_n27_dbg${2:-0}.txt
done
