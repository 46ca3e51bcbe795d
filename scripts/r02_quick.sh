# quick check on one B200: a pytest selection + the default workload without the extras.
# usage: bash scripts/r02_quick.sh TAG ["pytest -k expression"]
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-q}; K=${2:-"stream or vcycle or mg_8192 or 1024"}
timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 --no-also --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('default', round(d['ms_per_step'],4), '%.4e'%d['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), round(d['roofline']['ms_per_launch'],4))
print({k:v['ms'] for k,v in d['roofline']['stages'].items()})
PY
