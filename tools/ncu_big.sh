#!/bin/bash
# ncu --set full of the 416^2 tail of one trajectory-decoder chunk (upconv, ring fix, decoder.4.0, decoder.4.2, predictor)
# and of the TTST kernels.  Launch indices follow profiles/launches_*.csv (graph replay, profiler range = timed region).
# The reports stay in /tmp on the box (gpurun_out/ is capped at 64 MiB); the summaries and the tail report come back.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -s ${NCU_SKIP:-85} -c ${NCU_COUNT:-10} -f -o /tmp/prof_tail \
   python bench.py --steps 1 --warmup 3 --agents ${BENCH_AGENTS:-128} --no-cpu-baseline --no-roofline > gpurun_out/ncu_tail.log 2>&1; echo "ncu tail exit $?"
python tools/ncu_summary.py /tmp/prof_tail.ncu-rep > gpurun_out/ncu_full_tail416.md
for i in 0 4 6 7 9; do python tools/ncu_stalls.py /tmp/prof_tail.ncu-rep $i 12 > gpurun_out/ncu_stalls_tail_$i.txt 2>&1; done
timeout 900 ncu --set full --clock-control none --profile-from-start off \
   -s 42 -c 20 -f -o /tmp/prof_ttst \
   python bench.py --steps 1 --warmup 3 --agents ${BENCH_AGENTS:-128} --no-cpu-baseline --no-roofline > gpurun_out/ncu_ttst.log 2>&1; echo "ncu ttst exit $?"
python tools/ncu_summary.py /tmp/prof_ttst.ncu-rep > gpurun_out/ncu_full_ttst.md
ls -la /tmp/*.ncu-rep
[ $(stat -c %s /tmp/prof_tail.ncu-rep) -lt 45000000 ] && cp /tmp/prof_tail.ncu-rep gpurun_out/
ls -la gpurun_out
