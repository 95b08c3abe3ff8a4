#!/bin/bash
# fp32-parity family (bf16x3, mma.sync kernels): launch list, then ncu --set full of its edge kernels
mkdir -p gpurun_out
CMD="python tools/prof_step.py --precision bf16x3 --mols 5000 --fixed-atoms 0 --steps 2"
$CMD > gpurun_out/ncu_x3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_x3_launches.csv $CMD > gpurun_out/ncu_x3_list.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'edge_kernel|node_mlp_kernel' -s 10 -c 6 -o gpurun_out/r2_prof_x3 $CMD > gpurun_out/ncu_x3_full.log 2>&1
tail -3 gpurun_out/ncu_x3_plain.log; tail -1 gpurun_out/ncu_x3_full.log
