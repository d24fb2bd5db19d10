#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_kmeans.py 128 2>&1 | tail -2 | tee gpurun_out/kmeans_bench.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kmeans_kernel -s 2 -c 1 -o gpurun_out/kmeans_r02 -f python tools/bench_kmeans.py 128 > gpurun_out/ncu_kmeans.log 2>&1; tail -2 gpurun_out/ncu_kmeans.log
