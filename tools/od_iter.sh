#!/bin/bash
# One k_orient_desc iteration on the GPU box: extraction parity tests, short device-resident bench per variant (stage times), then the
# scheduler / warp-state sections of ONE ncu pass over the batched kernel.  usage: bash tools/od_iter.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_golden.py tests/test_cpp_shim.py -m gpu -x -q > gpurun_out/pytest_od_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_od_$TAG.log
B="python bench.py --no-matching --cpu-seconds 0 --no-secondary --sustain-seconds 0"
for V in "ORBX_DESC_BATCH=0" "ORBX_DESC_RING=3" "ORBX_DESC_RING=4" "ORBX_DESC_RING=5"; do
  env $V timeout 300 $B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$V','value',round(d['value']),'e2e',round(d['e2e']['value']),'orient_desc',round(d['roofline']['stage_ms']['orient_desc'],4),'total',round(d['roofline']['stage_ms']['total'],4))"
done
timeout 600 ncu --clock-control none -k regex:k_orient_desc_batch -c 1 --section SchedulerStats --section WarpStateStats --section Occupancy --section LaunchStats --section InstructionStats \
  --log-file gpurun_out/ncu_od_$TAG.txt $B --steps 1 --warmup 1 --frames 64 > gpurun_out/ncu_od_$TAG.log 2>&1; echo "ncu rc=$?"
grep -E "Registers|Shared Memory Config|Theoretical Occ|Achieved Occ|Issued Warp|No Eligible|Eligible Warps|Stall|Executed Ipc|Warp Cycles Per Issued|Executed Instructions|Block Limit" gpurun_out/ncu_od_$TAG.txt | head -40
