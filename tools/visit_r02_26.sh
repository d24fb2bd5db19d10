set -x
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --mode evaluate --agents 1024 --steps 3 --warmup 1 2>&1 | grep "^{" | cut -c1-300
