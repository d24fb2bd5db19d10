set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 3 --profile-layers gpurun_out/layers_r02_final.json > gpurun_out/bench_r02_final.log 2>&1; echo "bench exit $?"
grep "^{" gpurun_out/bench_r02_final.log | cut -c1-200
