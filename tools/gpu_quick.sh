#!/bin/bash
# quick check after a kernel change: bf16 errors vs fixtures, then step / per-kernel timings on both bench shapes
mkdir -p gpurun_out
timeout 180 python tools/check_edge_impl.py > gpurun_out/chk.log 2>&1; echo "chk rc=$?" >> gpurun_out/chk.log; tail -8 gpurun_out/chk.log
timeout 300 python tools/prof_step.py --mols 16384 --fixed-atoms 27 > gpurun_out/prof27.log 2>&1; echo "rc=$?" >> gpurun_out/prof27.log; tail -12 gpurun_out/prof27.log
timeout 300 python tools/prof_step.py --mols 5000 --fixed-atoms 0 > gpurun_out/prof_prior.log 2>&1; echo "rc=$?" >> gpurun_out/prof_prior.log; tail -12 gpurun_out/prof_prior.log
