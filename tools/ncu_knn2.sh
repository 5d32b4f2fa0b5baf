#!/bin/bash
# ncu --set full capture of the shipped brute-force 2-NN kernels (k_knn2 + merge) at 100k x 100k; report -> gpurun_out/full_knn2_$1.ncu-rep
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_knn2" -c 2 -o gpurun_out/full_knn2_$TAG -f \
    python bench.py --steps 1 --warmup 1 --frames 8 --cpu-seconds 0 --no-secondary --sustain-seconds 0 > gpurun_out/ncu_knn2_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/full_knn2_$TAG.ncu-rep --page raw --csv > gpurun_out/knn2_raw_$TAG.csv 2>/dev/null
