#!/bin/bash
# full evidence run: smoke, all GPU parity tests, bench (both arms), ncu launch list + full capture of the member kernel
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 300 python tools/profile_run.py --members 148 --years 2 > gpurun_out/profile_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python tools/profile_run.py --members 148 --years 2 > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/profile_run.py --members 148 --years 2 > gpurun_out/profile_plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:greb_member -s 1 -c 1 -o gpurun_out/prof python tools/profile_run.py --members 148 --years 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
