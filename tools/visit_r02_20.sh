set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rowconv.py tests/test_gpu_parity_benched.py tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/pytest_r2.log 2>&1; tail -4 gpurun_out/pytest_r2.log | cut -c1-300
for G in 1 0; do YNET_ROWCONV2=$G timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 --profile-layers gpurun_out/layers_r02_rc2_$G.json > gpurun_out/bench_r02_rc2_$G.log 2>&1; grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"' gpurun_out/bench_r02_rc2_$G.log; done
YNET_ROWCONV2_TAIL=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 > gpurun_out/bench_r02_rc2_tail.log 2>&1; grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"' gpurun_out/bench_r02_rc2_tail.log
