#!/bin/bash
# timing ablations of the edge pipeline (-DSMB_DEBUG build; results are numerically wrong when a bit is set):
#   1 = no E2 body, 2 = no LN apply / z store, 4 = no P compute (after the first two tiles)
mkdir -p gpurun_out
for d in ${@:-0 1 2 4 7}; do
  echo "== SMB_WS_DBG=$d"
  SMB_WS_DBG=$d timeout 200 python tools/prof_step.py --mols 8192 --fixed-atoms 27 --steps 5 2>&1 | grep -E "edge_|step " | tee -a gpurun_out/ablate.log
done
