#!/bin/bash
# One GPU-box visit (round 2): GPU parity tests, smoke, the default bench line (configs[2] + secondary legs), the reference arm.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-1500
