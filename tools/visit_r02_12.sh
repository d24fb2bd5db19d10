set -x
timeout 600 python -m pytest tests/test_gpu_cws_edge.py -q -m gpu > gpurun_out/pytest_cws.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_cws.log; tail -n 15 gpurun_out/pytest_cws.log
timeout 600 python bench.py --workload ind_short_ynetmod --steps 10 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 > gpurun_out/bench_r02_ynetmod.log 2>&1; tail -c 300 gpurun_out/bench_r02_ynetmod.log
for d in 0 1 2 3 4 7; do echo "DBG $d"; YNET_RC_DBG=$d N=320 MODE=fused python tools/bench_rowconv.py 2>&1 | tail -n 1; done
