#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_scripts.py tests/test_gpu_network.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
for B in bf16x3 fp32; do timeout 600 python bench.py --mode finetune --workload sdd_short --agents 30 --steps 10 --warmup 3 --backend $B > gpurun_out/bench_r02_finetune_sdd_1gpu_$B.log 2>&1; grep "^{" gpurun_out/bench_r02_finetune_sdd_1gpu_$B.log | tail -n 1 | cut -c1-400; done
timeout 600 python bench.py --mode finetune --workload ind_short_ynetmod --agents 30 --steps 10 --warmup 3 --backend bf16x3 > gpurun_out/bench_r02_finetune_ynetmod_1gpu_bf16x3.log 2>&1; grep "^{" gpurun_out/bench_r02_finetune_ynetmod_1gpu_bf16x3.log | tail -n 1 | cut -c1-300
HOST_PROFILE=1 timeout 500 python tools/profile_finetune.py bf16x3 2>&1 | grep "host issue"
