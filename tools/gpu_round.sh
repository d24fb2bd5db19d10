#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
# tensor-core engine first, in its own process (a protocol bug traps instead of hanging; keep it isolated)
timeout 400 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
tail -30 gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --ignore tests/test_gpu_tc.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-2} --warmup ${BENCH_WARMUP:-1} --agents ${BENCH_AGENTS:-16} \
   --profile-layers gpurun_out/layers.json ${BENCH_EXTRA} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
