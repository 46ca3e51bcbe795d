# round 2, one B200: full GPU suite, smoke, the default bench line (with its `also` list) and the reference arm
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-r02a}
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -14 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 600 gpurun_out/${TAG}_bench_reference.json; echo; grep -ci "non-finite" gpurun_out/${TAG}_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -5 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_default.json').read().strip().splitlines()[-1])
print('default', round(d['ms_per_step'],4), '%.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), 'cpu', d.get('cpu_baseline'))
print({k:v['ms'] for k,v in d['roofline']['stages'].items()})
for a in d.get('also',[]): print(a['workload'], round(a['ms_per_step'],4), '%.4e'%a['value'], a['roofline']['kernel'], round(a['roofline']['frac'],3), 'whole', round(a['whole_step_frac'],3))
PY
free -g | head -2; nproc
