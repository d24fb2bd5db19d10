set -x
timeout 1200 python -m pytest tests/test_gpu_rowconv.py -q -x -m gpu > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tc.log
tail -n 3 gpurun_out/pytest_tc.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-layers gpurun_out/layers_r02e.json > gpurun_out/bench_r02e.log 2>&1
grep -o '"value": [0-9.]*, "unit": "agent-trajectories/s", "n_gpus"' gpurun_out/bench_r02e.log
N=160 MODE=fused REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_rowconv -s 3 -c 1 -f -o gpurun_out/prof_row_fused python tools/bench_rowconv.py > gpurun_out/ncu_row_fused.log 2>&1
