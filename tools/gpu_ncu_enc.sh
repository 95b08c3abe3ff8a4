#!/bin/bash
# ncu launch list of the VN-DGCNN encoder (256 clouds x 1024 points)
mkdir -p gpurun_out
CMD="python tools/bench_encoder.py --clouds 256 --chunk 256"
$CMD > gpurun_out/ncu_enc_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_enc_launches.csv $CMD > gpurun_out/ncu_enc_list.log 2>&1
tail -1 gpurun_out/ncu_enc_plain.log | cut -c1-200
