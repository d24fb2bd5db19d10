set -x
timeout 600 python -m pytest tests/test_gpu_rowconv.py -q -m gpu > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tc.log; tail -n 3 gpurun_out/pytest_tc.log
for d in 0 7; do echo "DBG $d"; YNET_RC_DBG=$d N=320 MODE=fused python tools/bench_rowconv.py 2>&1 | tail -n 1; done
N=320 python tools/bench_rowconv.py 2>&1 | tail -n 5
