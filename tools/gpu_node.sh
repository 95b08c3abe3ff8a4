#!/bin/bash
mkdir -p gpurun_out
SMB_NODE_LEGACY=1 timeout 300 python tools/check_edge_impl.py > gpurun_out/chk_ws_nl.log 2>&1; echo rc $?
timeout 300 python tools/check_edge_impl.py > gpurun_out/chk_ws.log 2>&1; echo rc $?
cat gpurun_out/chk_ws_nl.log gpurun_out/chk_ws.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/bench_roles.sh "node_pre edge_k"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I shapemol_b200/csrc tools/tmem_bw_probe.cu -o /tmp/tmem_probe && timeout 60 /tmp/tmem_probe
