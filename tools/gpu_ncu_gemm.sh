#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bench_generic.py"
$CMD > gpurun_out/ncu_gen_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_persistent -s 4 -c 2 -o gpurun_out/r2_prof_gemm $CMD > gpurun_out/ncu_gemm_full.log 2>&1
tail -2 gpurun_out/ncu_gemm_full.log
