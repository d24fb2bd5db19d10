#!/bin/bash
# Round-end evidence: all GPU tests, smoke, the default bench line, launch list + DRAM traffic of one step (ncu metrics
# pass), ncu --set full of the 416^2 tail.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   --profile-from-start off -c 900 --csv --log-file gpurun_out/traffic.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
timeout 900 python bench.py --profile-layers gpurun_out/layers.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -c 3000 gpurun_out/bench.log
