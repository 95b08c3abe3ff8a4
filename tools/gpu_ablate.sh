#!/bin/bash
# timing ablations of the warp-specialised edge pipeline (results are wrong with SMB_WS_DBG set; timing only)
for d in 0 1 2 4 3 5 6 7; do
  for k in edge_k edge_v edge_xv; do
    SMB_WS_DBG=$d timeout 200 python bench.py --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-mode --prof-kernel $k > /tmp/ab.log 2>&1
    python - <<PY
import json
l=[x for x in open("/tmp/ab.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("dbg $d $k ms_per_launch %.4f" % d["roofline"]["ms_per_launch"])
else:
    print("dbg $d $k failed")
PY
  done
done
