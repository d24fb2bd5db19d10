#!/bin/bash
# N-GPU confirmation at HEAD: NCCL tests (incl. the train / test entry points under torchrun), default and evaluate lines.
N=$1
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR -m pytest tests/test_gpu_nccl.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_nccl_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nccl_${N}gpu.log
tail -n 12 gpurun_out/pytest_nccl_${N}gpu.log
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-roofline > gpurun_out/bench_r02_default_${N}gpu.log 2>&1
timeout 600 $TR bench.py --gpus $N --mode evaluate --agents 1024 --steps 3 --warmup 1 > gpurun_out/bench_r02_evaluate_${N}gpu.log 2>&1
for f in gpurun_out/bench_r02_default_${N}gpu.log gpurun_out/bench_r02_evaluate_${N}gpu.log; do echo "== $f"; grep "^{" $f | tail -n 1 | cut -c1-330; done
