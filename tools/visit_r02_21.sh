set -x
timeout 600 python -m pytest tests/test_gpu_rowconv.py -x -q -m gpu -k rowconv2 2>&1 | tail -2
for D in 0 1; do echo "DBG $D"; YNET_RC_DBG=$D MODE=rc2 N=320 timeout 300 python tools/bench_rowconv.py 2>&1 | grep "two-conv\|L1 row"; done
