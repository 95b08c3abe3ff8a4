#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_edge_impl.py 2>&1 | tail -8
bash tools/bench_roles.sh "edge_k edge_v edge_xv node_pre"
echo "--- suspend hint 1000 ns"
SMB_NVCC_EXTRA=-DSMB_MBAR_SUSPEND_HINT=1000 python shapemol_b200/build.py --force > /dev/null 2>&1; echo build rc $?
timeout 300 python tools/check_edge_impl.py 2>&1 | tail -2
bash tools/bench_roles.sh "edge_k edge_v edge_xv node_pre"
