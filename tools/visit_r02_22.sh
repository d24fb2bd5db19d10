set -x
mkdir -p gpurun_out
MODE=rc2 N=160 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_rowconv2_kernel -s 3 -c 1 -f -o /tmp/prof_rc2_plain python tools/bench_rowconv.py > gpurun_out/ncu_rc2_plain.log 2>&1
MODE=rc2 N=160 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_rowconv2_kernel -s 7 -c 1 -f -o /tmp/prof_rc2_tail python tools/bench_rowconv.py > gpurun_out/ncu_rc2_tail.log 2>&1
for v in plain tail; do
  python tools/ncu_summary.py /tmp/prof_rc2_$v.ncu-rep > gpurun_out/ncu_r02_rowconv2_$v.md 2>&1
  ncu -i /tmp/prof_rc2_$v.ncu-rep --page source --csv --print-source cuda > gpurun_out/ncu_rc2_${v}_source_cuda.csv 2>&1
  python tools/ncu_stalls.py /tmp/prof_rc2_$v.ncu-rep 0 40 > gpurun_out/ncu_rc2_${v}_stalls.txt 2>&1
done
ls -la gpurun_out/ncu_rc2* /tmp/*.ncu-rep
