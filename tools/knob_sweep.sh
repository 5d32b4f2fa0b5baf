#!/bin/bash
# Short device-resident bench per environment setting: usage: bash tools/knob_sweep.sh "VAR=a" "VAR=b" ...
B="python bench.py --no-matching --cpu-seconds 0 --no-secondary --sustain-seconds 0"
for V in "$@"; do
  env $V timeout 300 $B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['roofline']['stage_ms']
print('$V','value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'stages',{k:round(v,3) for k,v in s.items()})"
done
