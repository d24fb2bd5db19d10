set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-200
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log; tail -n 4 gpurun_out/sanitize_$tool.log | cut -c1-200
done
