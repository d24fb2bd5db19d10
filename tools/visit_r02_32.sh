#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_kmeans.py 128 2>&1 | tail -5 | tee gpurun_out/kmeans_variants.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k kmeans 2>&1 | tail -3
