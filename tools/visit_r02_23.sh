set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/pytest_traintc.log 2>&1; tail -5 gpurun_out/pytest_traintc.log | cut -c1-300
for B in fp32 bf16x3; do timeout 600 python bench.py --mode finetune --workload sdd_short --agents 30 --steps 3 --warmup 1 --backend $B > gpurun_out/bench_r02_finetune_sdd_1gpu_$B.log 2>&1; grep "^{" gpurun_out/bench_r02_finetune_sdd_1gpu_$B.log | cut -c1-260; done
timeout 600 python bench.py --mode finetune --workload ind_short_ynetmod --agents 30 --steps 3 --warmup 1 --backend bf16x3 > gpurun_out/bench_r02_finetune_ynetmod_1gpu_bf16x3.log 2>&1; grep "^{" gpurun_out/bench_r02_finetune_ynetmod_1gpu_bf16x3.log | cut -c1-260
