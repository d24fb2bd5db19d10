set -x
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_cws_edge.py -q -m gpu > gpurun_out/pytest_pre.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pre.log; tail -n 25 gpurun_out/pytest_pre.log
