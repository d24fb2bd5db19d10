set -x
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -2
for MB in 1 128; do YNET_EVAL_MIN_BATCH=$MB timeout 600 python bench.py --mode evaluate --agents 1024 --chunk-agents 10 --steps 3 --warmup 1 2>&1 | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('min_batch', '$MB', d['value'], d['ms_per_step'], d['config'].get('batch_size'))"; done
