# 1-GPU measurement campaign of round 2: every bench line the verdict asked for (profiles/bench_r02_*.json)
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --profile-layers gpurun_out/layers_r02.json > gpurun_out/bench_r02_default.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.log 2>&1
timeout 600 python bench.py --workload sdd_short --steps 10 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 > gpurun_out/bench_r02_sdd_short.log 2>&1
timeout 600 python bench.py --workload ind_short_ynetmod --steps 10 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 > gpurun_out/bench_r02_ynetmod.log 2>&1
timeout 900 python bench.py --backend fp32 --agents 32 --steps 3 --warmup 3 --no-cpu-baseline --torch-cuda-agents 0 --no-graph > gpurun_out/bench_r02_fp32.log 2>&1
timeout 600 python bench.py --agents 1024 --chunk-agents 128 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r02_sweep_1k.log 2>&1
timeout 600 python bench.py --agents 8192 --chunk-agents 128 --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r02_sweep_8k.log 2>&1
timeout 900 python bench.py --agents 65536 --chunk-agents 128 --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r02_sweep_64k.log 2>&1
timeout 600 python bench.py --mode evaluate --agents 1024 --steps 2 --warmup 1 > gpurun_out/bench_r02_evaluate_1gpu.log 2>&1
timeout 600 python bench.py --mode finetune --workload sdd_short --agents 30 --steps 3 --warmup 1 > gpurun_out/bench_r02_finetune_sdd_1gpu.log 2>&1
timeout 600 python bench.py --mode finetune --workload ind_short_ynetmod --agents 30 --steps 3 --warmup 1 > gpurun_out/bench_r02_finetune_ynetmod_1gpu.log 2>&1
for f in gpurun_out/bench_r02_*.log; do echo "== $f"; tail -c 400 $f | tail -n 2 | cut -c1-400; done
