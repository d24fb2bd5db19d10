set -x
N=160 MODE=fused REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_rowconv -s 3 -c 1 -f -o gpurun_out/prof_row_fused python tools/bench_rowconv.py > gpurun_out/ncu_row_fused.log 2>&1
N=160 MODE=plain REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_rowconv -s 3 -c 1 -f -o gpurun_out/prof_row_plain python tools/bench_rowconv.py > gpurun_out/ncu_row_plain.log 2>&1
ls -la gpurun_out/*.ncu-rep
