set -x
for D in 0 8 16 24 31; do echo "DBG $D"; YNET_RC_DBG=$D MODE=rc2 N=320 timeout 300 python tools/bench_rowconv.py 2>&1 | grep "two-conv"; done
