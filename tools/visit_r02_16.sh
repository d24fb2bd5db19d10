set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_train.py -x -q -m gpu -s > gpurun_out/pytest_split.log 2>&1; grep -n "split conv\|bf16x3\|passed\|failed\|Error" gpurun_out/pytest_split.log | head -40
timeout 900 python -m pytest tests/test_gpu_parity_benched.py -x -q -m gpu -s -k bf16x3 > gpurun_out/pytest_split_parity.log 2>&1; grep -n "bf16x3\|passed\|failed\|Error" gpurun_out/pytest_split_parity.log | head -20
timeout 900 python bench.py --backend bf16x3 --steps 5 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 --profile-layers gpurun_out/layers_r02_bf16x3.json > gpurun_out/bench_r02_bf16x3.log 2>&1; grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"\|"allocated_peak_gb": [0-9.]*\|"dtype": "[a-z0-9]*"' gpurun_out/bench_r02_bf16x3.log | head
