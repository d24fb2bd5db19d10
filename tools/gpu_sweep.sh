#!/bin/bash
# Chunk-size sweep of the trajectory decoder + the default bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_network.py -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log; tail -3 gpurun_out/pytest_quick.log
for P in 256 320 640; do
  YNET_MAX_STACKED_PASSES=$P timeout 300 python bench.py --steps 8 --warmup 3 --agents 64 --no-cpu-baseline --no-roofline > gpurun_out/bench_p$P.log 2>&1
  echo "passes $P: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_p$P.log | head -1)"
done
YNET_MAX_STACKED_PASSES=640 timeout 400 python bench.py --steps 6 --warmup 3 --agents 128 --no-cpu-baseline --no-roofline > gpurun_out/bench_a128_p640.log 2>&1
echo "agents 128 passes 640: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_a128_p640.log | head -1)"
timeout 400 python bench.py --steps 6 --warmup 3 --agents 128 --no-cpu-baseline --no-roofline > gpurun_out/bench_a128.log 2>&1
echo "agents 128 passes 256: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_a128.log | head -1)"
