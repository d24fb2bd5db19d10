set -x
timeout 900 python -m pytest tests/test_gpu_rowconv.py -q -x -m gpu > gpurun_out/pytest_rowconv.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_rowconv.log
tail -n 30 gpurun_out/pytest_rowconv.log
if grep -q "rc=0" gpurun_out/pytest_rowconv.log; then
  timeout 1500 python -m pytest tests/test_gpu_parity_benched.py -q -s -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity.log
  grep -e "^\.*\[" -e passed -e failed gpurun_out/pytest_parity.log | cut -c1-300
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-layers gpurun_out/layers_r02c.json > gpurun_out/bench_r02c.log 2>&1
  YNET_ROWCONV=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r02c_off.log 2>&1
  grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"' gpurun_out/bench_r02c.log gpurun_out/bench_r02c_off.log
fi
