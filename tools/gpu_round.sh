#!/bin/bash
# One GPU round: gpu tests, smoke, bench and the reference arm; with NCU=launches or NCU=full ONLY the corresponding profiler pass
# (ncu launch list / ncu --set full of the dominant kernels) — numbers measured under a profiler are never bench values.
# usage (on the GPU box, from the repo root): [NCU=launches|full] bash tools/gpu_round.sh <tag> [kernel-regex for the full capture]
TAG=${1:-r1}
KREGEX=${2:-"k_fast|k_blur|k_resize|k_knn2|k_octree|k_orient_desc|k_level0"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.log 2>&1
if [ -z "$NCU" ] || [ "$NCU" = "0" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"; cat gpurun_out/bench_ref_$TAG.log
fi
# profiler passes: one per call (NCU=launches or NCU=full), never together with the timed runs above
if [ "$NCU" = "launches" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --frames 64 --knn-n 100000 --cpu-seconds 0 --no-secondary --sustain-seconds 0 > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
elif [ "$NCU" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -c 30 -o gpurun_out/full_$TAG -f \
    python bench.py --steps 1 --warmup 1 --frames 64 --knn-n 100000 --cpu-seconds 0 --no-secondary --sustain-seconds 0 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
