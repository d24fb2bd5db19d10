set -x
N=160 MODE=fused REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_rowconv -s 3 -c 1 -f -o gpurun_out/prof_row_fused python tools/bench_rowconv.py > gpurun_out/ncu_row_fused.log 2>&1
N=160 MODE=l2 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_rowconv -s 3 -c 1 -f -o gpurun_out/prof_row_l2 python tools/bench_rowconv.py > gpurun_out/ncu_row_l2.log 2>&1
