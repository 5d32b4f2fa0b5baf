#!/bin/bash
# On the GPU box: for every variants/liborbx_<name>.so given, install it as the product library, run the extraction parity tests and
# a short device-resident bench, print the stage times.  The original library is restored at the end.
# usage: bash tools/variant_bench.sh <name>...   ("base" = the library as built)
mkdir -p gpurun_out
cp openmavis_b200/liborbx.so /tmp/liborbx_base.so
B="python bench.py --no-matching --cpu-seconds 0 --no-secondary --sustain-seconds 0"
for V in "$@"; do
  if [ "$V" = base ]; then cp /tmp/liborbx_base.so openmavis_b200/liborbx.so; else cp variants/liborbx_$V.so openmavis_b200/liborbx.so; fi
  if [ -z "$NOTEST" ]; then timeout 600 python -m pytest tests/test_gpu_extract.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2; fi
  for rep in 1 2; do
  timeout 300 $B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['roofline']['stage_ms']
print('$V','value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'stages',{k:round(v,4) for k,v in s.items()})"
  done
done
cp /tmp/liborbx_base.so openmavis_b200/liborbx.so
