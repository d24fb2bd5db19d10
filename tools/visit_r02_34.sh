#!/bin/bash
# HEAD confirmation after the k-means change and the scripts' host side: all GPU tests, smoke, sanitizer, default bench,
# evaluate() drop-in bench.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log; tail -n 3 gpurun_out/sanitize_$tool.log
done
timeout 900 python bench.py --steps 20 --warmup 3 --profile-layers gpurun_out/layers_r02_final.json > gpurun_out/bench_r02_final.log 2>&1; echo "bench exit $?"
grep "^{" gpurun_out/bench_r02_final.log | tail -n 1 | cut -c1-400
timeout 600 python bench.py --mode evaluate --agents 1024 --steps 3 --warmup 1 > gpurun_out/bench_r02_evaluate_1gpu.log 2>&1
grep "^{" gpurun_out/bench_r02_evaluate_1gpu.log | tail -n 1 | cut -c1-300
