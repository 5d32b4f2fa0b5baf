#!/bin/bash
# One k_fast iteration on the GPU box: extraction parity tests, a short device-resident bench (stage times + FAST counters), then ONE
# ncu pass over k_fast (instruction counts / pipes) of a separate run.  usage: bash tools/fast_iter.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_golden.py tests/test_cpp_shim.py -m gpu -x -q > gpurun_out/pytest_fast_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_fast_$TAG.log
timeout 300 python bench.py --no-matching --cpu-seconds 0 --no-secondary --sustain-seconds 0 > gpurun_out/bench_fast_$TAG.log 2>gpurun_out/bench_fast_$TAG.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_fast_$TAG.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e',d['e2e']['value'])
print(d['roofline']['stage_ms'])
print(d['run'].get('fast_counters'))"
timeout 600 ncu --clock-control none -k regex:k_fast -c 2 --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,dram__bytes_read.sum \
  --log-file gpurun_out/ncu_fast_$TAG.csv python bench.py --steps 1 --warmup 1 --frames 64 --no-matching --cpu-seconds 0 --no-secondary --sustain-seconds 0 > gpurun_out/ncu_fast_$TAG.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_fast_$TAG.csv')) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r))
    if d.get('ID')=='0' or True:
        print(d['Metric Name'], d['Metric Value'])
PY
