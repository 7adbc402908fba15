#!/bin/bash
# quick loop: fast parity tests, profile workload timing, short bench, ncu summary of the member kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_run.py --members 148 --years 2 > gpurun_out/profile_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:greb_member -s 1 -c 1 -o gpurun_out/prof python tools/profile_run.py --members 148 --years 2 > gpurun_out/ncu_full.log 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
