set -x
timeout 1500 python -m pytest tests/test_gpu_parity_benched.py -q -s -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity.log
YNET_HOIST_LO=1 timeout 1500 python -m pytest tests/test_gpu_parity_benched.py -q -s -m gpu -k "oracle_waypoints" > gpurun_out/pytest_parity_hoistlo.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-roofline > gpurun_out/bench_r02b.log 2>&1
tail -n 3 gpurun_out/pytest_parity.log
