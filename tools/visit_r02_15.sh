set -x
for P in 640 1280 2560; do echo "PASSES $P"; YNET_MAX_STACKED_PASSES=$P timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --torch-cuda-agents 0 2>&1 | grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"\|"allocated_peak_gb": [0-9.]*' | head -2; done
echo "AGENTS 256"; YNET_MAX_STACKED_PASSES=2560 timeout 600 python bench.py --agents 256 --steps 6 --warmup 3 --no-cpu-baseline --no-roofline --torch-cuda-agents 0 2>&1 | grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"\|"allocated_peak_gb": [0-9.]*' | head -2
