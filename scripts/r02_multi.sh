# round 2, N B200 of one box: the row-slab suite (N-slab runs against one GPU, both transports) and bench.py --gpus N
# with its parity block.  usage: bash scripts/r02_multi.sh N TAG [notests]
cd /root/repo; mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02a}
if [ "$3" != "notests" ]; then
  timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${TAG}_pytest_multi${N}.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_multi${N}.log
fi
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err ) 2>&1 | grep real; tail -3 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n${N}.json').read().strip().splitlines()[-1])
print('default N=$N', round(d['ms_per_step'],4), '%.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3))
print('parity', d.get('parity'))
print({k:v['ms'] for k,v in d['roofline']['stages'].items()})
for a in d.get('also',[]): print(a['workload'], a['decomposition'], round(a['ms_per_step'],4), '%.4e'%a['value'], round(a['roofline']['frac'],3), 'parity', a.get('parity'))
PY
