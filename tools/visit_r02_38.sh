#!/bin/bash
# final HEAD confirmation: all GPU tests, smoke, default bench line, bf16x3 fine-tune line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_final.log 2>&1; echo "bench exit $?"
grep "^{" gpurun_out/bench_r02_final.log | tail -n 1 | cut -c1-260
for B in fp32 bf16x3; do timeout 600 python bench.py --mode finetune --workload sdd_short --agents 30 --steps 20 --warmup 3 --backend $B > gpurun_out/bench_r02_finetune_sdd_1gpu_$B.log 2>&1; grep "^{" gpurun_out/bench_r02_finetune_sdd_1gpu_$B.log | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$B', round(d['ms_per_step'],2), [round(x) for x in d['step_ms']], d['clocks'])"; done
