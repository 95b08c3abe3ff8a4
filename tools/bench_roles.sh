#!/bin/bash
# per-kernel timing breakdown of one denoising step (bf16 mode): tools/bench_roles.sh [kernels...]
for k in "${@:-edge_k edge_v edge_xv node_pre}"; do for kk in $k; do
timeout 100 python bench.py --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-parity-mode --prof-kernel $kk > gpurun_out/b_ws_$kk.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/b_ws_$kk.log") if x.startswith("{")]
d=json.loads(l[-1]); print("$kk", round(d["ms_per_step"],3), round(d["mol_steps_per_s"]), {k:v for k,v in d["roofline"].items() if k in ("ms_per_launch","launches_per_step","share_of_step","frac")})
PY
done; done
