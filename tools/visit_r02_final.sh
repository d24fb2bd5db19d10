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
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.log 2>&1; echo "reference exit $?"
timeout 900 python bench.py --backend bf16x3 --steps 5 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 > gpurun_out/bench_r02_bf16x3.log 2>&1; echo "bf16x3 exit $?"
timeout 600 python bench.py --mode evaluate --agents 1024 --steps 3 --warmup 1 > gpurun_out/bench_r02_evaluate_1gpu.log 2>&1
for B in fp32 bf16x3; do timeout 600 python bench.py --mode finetune --workload sdd_short --agents 30 --steps 3 --warmup 2 --backend $B > gpurun_out/bench_r02_finetune_sdd_1gpu_$B.log 2>&1; done
for f in gpurun_out/bench_r02_reference.log gpurun_out/bench_r02_bf16x3.log gpurun_out/bench_r02_evaluate_1gpu.log gpurun_out/bench_r02_finetune_sdd_1gpu_fp32.log gpurun_out/bench_r02_finetune_sdd_1gpu_bf16x3.log; do echo "== $f"; grep "^{" $f | tail -n 1 | cut -c1-300; done
