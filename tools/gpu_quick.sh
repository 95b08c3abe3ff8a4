#!/bin/bash
# quick check: bf16 errors vs fixtures + per-role timings
mkdir -p gpurun_out
timeout 120 python tools/check_edge_impl.py 2>&1 | tail -7
bash tools/bench_roles.sh "${1:-edge_k edge_v edge_xv}"
