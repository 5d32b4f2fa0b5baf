#!/bin/bash
# ncu --set full capture of ONE kernel (regex $1, tag $2, launches $3) of a short device-resident bench run; report -> gpurun_out/full_$2.ncu-rep
K=${1:-k_fast}; TAG=${2:-x}; N=${3:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -c $N -o gpurun_out/full_$TAG -f \
    python bench.py --steps 1 --warmup 1 --frames 64 --no-matching --cpu-seconds 0 --no-secondary --sustain-seconds 0 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full_$TAG.log
