#!/bin/bash
# ncu --set full of the encoder's three heaviest kernels (one launch each), after the same command ran without ncu
mkdir -p gpurun_out
CMD="python tools/bench_encoder.py --clouds 256 --chunk 256"
$CMD > gpurun_out/ncu_enc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'enc_edge_apply_kernel|enc_topk_kernel|enc_edge_stats_kernel' -s 3 -c 3 -o gpurun_out/r2_prof_enc $CMD > gpurun_out/ncu_enc_full.log 2>&1
tail -2 gpurun_out/ncu_enc_full.log
