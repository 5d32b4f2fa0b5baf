# usage: bash tools/run_bench_n.sh <N> <tag>
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.log 2>gpurun_out/bench_${N}gpu_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu_$TAG.log').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'],'copy_roofline',d['e2e']['copy_roofline']['value'],'frac',d['e2e']['frac_of_copy_roofline'])
print('run',d['run']); print('sustained',d['sustained'])
print({k:v for k,v in d['matching'].items() if k!='guided'})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_${N}gpu_$TAG.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_${N}gpu_$TAG.log | cut -c1-300
