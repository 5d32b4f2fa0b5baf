set -x
timeout 600 python -m pytest tests/test_gpu_extract.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --no-matching --cpu-seconds 0 --no-secondary --sustain-seconds 0 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e',d['e2e']['value'])
print(d['roofline']['stage_ms'])"
