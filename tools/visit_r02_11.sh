set -x
timeout 1200 python -m pytest tests/test_gpu_rowconv.py tests/test_gpu_tc.py tests/test_gpu_parity_benched.py -q -m gpu > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tc.log
tail -n 4 gpurun_out/pytest_tc.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r02g.log 2>&1
grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"' gpurun_out/bench_r02g.log
