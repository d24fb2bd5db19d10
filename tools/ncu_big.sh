#!/bin/bash
# ncu --set full of the 416^2 tail of one trajectory-decoder chunk (upconv, ring fix, decoder.4.0, decoder.4.2, predictor)
# and of the TTST kernels.  Launch indices follow profiles/launches_*.csv (graph replay, profiler range = timed region).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -s ${NCU_SKIP:-85} -c ${NCU_COUNT:-10} -f -o gpurun_out/prof_tail \
   python bench.py --steps 1 --warmup 3 --agents ${BENCH_AGENTS:-64} --no-cpu-baseline --no-roofline > gpurun_out/ncu_tail.log 2>&1; echo "ncu tail exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -s 42 -c 20 -f -o gpurun_out/prof_ttst \
   python bench.py --steps 1 --warmup 3 --agents ${BENCH_AGENTS:-64} --no-cpu-baseline --no-roofline > gpurun_out/ncu_ttst.log 2>&1; echo "ncu ttst exit $?"
ls -la gpurun_out
