mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'copy_roofline',d['e2e']['copy_roofline']['value'],'frac',d['e2e']['frac_of_copy_roofline'])
print('stage',d['roofline']['stage_ms'])
print('sustained',d['sustained'])
print('cpu',json.dumps(d['cpu_baseline'],indent=1))
print('secondary',json.dumps(d['secondary'],indent=1))
print('fast',d['run']['fast_counters'])
g=d['matching']['guided']
for k,v in g.items(): print(k, json.dumps(v)[:900])
print({k:v for k,v in d['matching'].items() if k!='guided'})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"; cut -c1-700 gpurun_out/bench_ref_$TAG.log
