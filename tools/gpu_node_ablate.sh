#!/bin/bash
for d in ${DBGS:-0 7 15 23 31}; do for k in node_pre; do
SMB_NODE_DBG=$d timeout 200 python bench.py --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-mode --prof-kernel $k > /tmp/ab.log 2>&1
python - <<PY
import json
l=[x for x in open("/tmp/ab.log") if x.startswith("{")]
print("dbg $d $k", "%.4f" % json.loads(l[-1])["roofline"]["ms_per_launch"] if l else "failed")
PY
done; done
