set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowconv.py -x -q -m gpu -s > gpurun_out/pytest_wp.log 2>&1; tail -5 gpurun_out/pytest_wp.log
timeout 900 python -m pytest tests/test_gpu_parity_benched.py tests/test_gpu_tc.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/pytest_wp2.log 2>&1; tail -5 gpurun_out/pytest_wp2.log
for G in 1 0; do YNET_WP_GATHER=$G timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 --profile-layers gpurun_out/layers_r02_wp$G.json > gpurun_out/bench_r02_wp$G.log 2>&1; grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"' gpurun_out/bench_r02_wp$G.log; done
