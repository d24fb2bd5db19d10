set -x
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_split.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python tools/profile_finetune.py bf16x3 2>&1 | grep "split_unpack\|split_pack\|Self C\|tc_conv_kernel<3\|wgrad"
timeout 600 python bench.py --mode finetune --workload sdd_short --agents 30 --steps 3 --warmup 2 --backend bf16x3 2>&1 | grep "^{" | cut -c1-250
