# multi-GPU measurements of round 2: NCCL tests of the data-parallel drop-ins, weak-scaling default line, strong-scaling
# sharded evaluate(), fine-tune steps with the flat-gradient all-reduce.   usage: bash tools/visit_r02_multi.sh N
N=$1
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR -m pytest tests/test_gpu_nccl.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_nccl_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nccl_${N}gpu.log
tail -n 6 gpurun_out/pytest_nccl_${N}gpu.log
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-roofline > gpurun_out/bench_r02_default_${N}gpu.log 2>&1
timeout 600 $TR bench.py --gpus $N --mode evaluate --agents 1024 --steps 3 --warmup 1 > gpurun_out/bench_r02_evaluate_${N}gpu.log 2>&1
timeout 600 $TR bench.py --gpus $N --mode finetune --workload sdd_short --agents 30 --steps 3 --warmup 1 > gpurun_out/bench_r02_finetune_sdd_${N}gpu.log 2>&1
timeout 600 $TR bench.py --gpus $N --mode finetune --workload ind_short_ynetmod --agents 30 --steps 3 --warmup 1 > gpurun_out/bench_r02_finetune_ynetmod_${N}gpu.log 2>&1
timeout 600 $TR bench.py --gpus $N --mode finetune --workload sdd_short --agents 30 --steps 3 --warmup 1 --backend bf16x3 > gpurun_out/bench_r02_finetune_sdd_bf16x3_${N}gpu.log 2>&1
for f in gpurun_out/bench_r02_*_${N}gpu.log; do echo "== $f"; grep "^{" $f | tail -n 1 | cut -c1-330; done
