#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bench_generic.py"
$CMD > gpurun_out/ncu_gen_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_generic_launches.csv $CMD > gpurun_out/ncu_gen_list.log 2>&1
tail -1 gpurun_out/ncu_gen_plain.log | cut -c1-200
