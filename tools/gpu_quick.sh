#!/bin/bash
# Short GPU visit: the GPU parity tests + one bench line (+ per-layer table).  Env: BENCH_AGENTS, PYTEST_ARGS
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 ${PYTEST_ARGS} > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -15 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 --agents ${BENCH_AGENTS:-64} --no-cpu-baseline \
   --profile-layers gpurun_out/layers.json ${BENCH_EXTRA} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -c 2500 gpurun_out/bench.log
