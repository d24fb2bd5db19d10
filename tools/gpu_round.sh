#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, ncu launch list + one --set full capture.
# Everything lands in gpurun_out/.   Env: BENCH_AGENTS, BENCH_STEPS, SKIP_TESTS=1, SKIP_NCU=1, NCU_KERNEL=regex
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
# tensor-core engine first, in its own process (a protocol bug traps instead of hanging; keep it isolated)
timeout 400 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
tail -30 gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --ignore tests/test_gpu_tc.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
fi
timeout 600 python bench.py --steps ${BENCH_STEPS:-10} --warmup ${BENCH_WARMUP:-3} --agents ${BENCH_AGENTS:-64} \
   --profile-layers gpurun_out/layers.json ${BENCH_EXTRA} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
if [ -z "$SKIP_NCU" ]; then
# launch list of the timed region (graph replay: ncu sees each kernel node)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 700 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --agents ${BENCH_AGENTS:-64} --no-cpu-baseline --no-roofline \
   > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
# the top kernel, full set, three launches
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:${NCU_KERNEL:-tc_conv} -s ${NCU_SKIP:-30} -c ${NCU_COUNT:-4} -f -o gpurun_out/prof_top \
   python bench.py --steps 1 --warmup 3 --agents ${BENCH_AGENTS:-64} --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
fi
