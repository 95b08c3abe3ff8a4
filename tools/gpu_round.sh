#!/bin/bash
# one GPU-box visit: parity tests, bench (bf16), ncu launch list, full ncu capture of the edge K kernel, role trace
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_bf16.log 2>&1 || { tail -20 gpurun_out/bench_bf16.log; exit 1; }
tail -c 600 gpurun_out/bench_bf16.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'edge_ws_kernel|node_tc5_kernel' --launch-skip 60 --launch-count 6 \
  -o gpurun_out/prof_edge_ws -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_full.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 400 gpurun_out/bench_reference.log
for role in 1 2 3; do SMB_WS_DBG=$((16 + role * 256)) python tools/ws_trace.py > gpurun_out/trace_role$role.log 2>&1; done
tail -16 gpurun_out/trace_role1.log
