set -x
timeout 900 python -m pytest tests/test_gpu_rowconv.py tests/test_gpu_parity_benched.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 2>&1 | grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"'
