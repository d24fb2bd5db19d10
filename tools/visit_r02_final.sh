#!/bin/bash
# Round-2 evidence at HEAD: all GPU tests, smoke, sanitizer over the new kernels, launch list + DRAM traffic of one step,
# the default bench line.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 5 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log; tail -n 4 gpurun_out/sanitize_$tool.log
done
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   --profile-from-start off -c 900 --csv --log-file gpurun_out/traffic.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --torch-cuda-agents 0 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
timeout 900 python bench.py --steps 20 --warmup 3 --profile-layers gpurun_out/layers_r02_final.json > gpurun_out/bench_r02_final.log 2>&1; echo "bench exit $?"
tail -c 600 gpurun_out/bench_r02_final.log
