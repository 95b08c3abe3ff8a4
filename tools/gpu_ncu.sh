#!/bin/bash
# ncu evidence for profiles/: launch list (per-launch device times) and one --set full capture of the edge / node kernels.
# The same command runs without ncu first (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python bench.py --quick --steps 2 --warmup 3 --mols 5000 --fixed-atoms 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'edge_ws_kernel|node_tc5_kernel' -s 12 -c 7 -o gpurun_out/r2_prof_edge $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_list.log; tail -2 gpurun_out/ncu_full.log
