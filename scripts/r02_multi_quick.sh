# N B200 of one box, short: a selection of the row-slab suite and bench.py --gpus N without the extras (parity block kept).
# usage: bash scripts/r02_multi_quick.sh N TAG
cd /root/repo; mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02u}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "jet-mg-1-1-0 or shear-mg-1-1-0 or lid-mg-0-1-0 or jet-mg-1-0-0 or jet-mg-1-1-33 or lid-pcg-1-1-0 or (late_rank and shear) or (bottom_top and shear-mg) or jet-512-256" > gpurun_out/${TAG}_pytest_multi${N}.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_multi${N}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 12 --warmup 3 --no-e2e --no-cpu-baseline --no-also > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; tail -2 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n${N}.json').read().strip().splitlines()[-1])
print('default N=$N', round(d['ms_per_step'],4), '%.4e'%d['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3))
print('parity', d.get('parity'))
print({k:v['ms'] for k,v in d['roofline']['stages'].items()})
PY
