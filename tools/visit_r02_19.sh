set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -m gpu > gpurun_out/pytest_evalgraph.log 2>&1; tail -5 gpurun_out/pytest_evalgraph.log
timeout 600 python bench.py --mode evaluate --agents 1024 --steps 3 --warmup 1 > gpurun_out/bench_r02_evaluate_1gpu.log 2>&1; grep "^{" gpurun_out/bench_r02_evaluate_1gpu.log | cut -c1-400
YNET_EVAL_RNG=host timeout 600 python bench.py --mode evaluate --agents 1024 --steps 3 --warmup 1 > gpurun_out/bench_r02_evaluate_1gpu_host.log 2>&1; grep "^{" gpurun_out/bench_r02_evaluate_1gpu_host.log | cut -c1-400
