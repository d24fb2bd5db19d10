set -x
mkdir -p gpurun_out
# ncu --set full of the three row kernels of one bench step (graph replay, profiler range = timed region)
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_rowconv -c 12 -f -o /tmp/prof_final \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --torch-cuda-agents 0 > gpurun_out/ncu_final.log 2>&1; echo "ncu exit $?"
python tools/ncu_summary.py /tmp/prof_final.ncu-rep > gpurun_out/ncu_r02_final_rowkernels.md 2>&1
ls -la /tmp/prof_final.ncu-rep; grep -c "^## " gpurun_out/ncu_r02_final_rowkernels.md
