set -x
timeout 1200 python -m pytest tests/test_gpu_rowconv.py -q -x -m gpu > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tc.log
tail -n 3 gpurun_out/pytest_tc.log
N=320 python tools/bench_rowconv.py > gpurun_out/bench_rowconv.log 2>&1; cat gpurun_out/bench_rowconv.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-layers gpurun_out/layers_r02f.json > gpurun_out/bench_r02f.log 2>&1
grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"' gpurun_out/bench_r02f.log
