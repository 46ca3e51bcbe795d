# round 2, one B200: the reference arm and the default bench line only
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-r02a}
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err ) 2>&1 | grep real; tail -c 700 gpurun_out/${TAG}_bench_reference.json; echo; echo "non-finite lines in reference stderr:"; grep -ci "non-finite" gpurun_out/${TAG}_bench_reference.err
( time python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err ) 2>&1 | grep real; tail -5 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_default.json').read().strip().splitlines()[-1])
print('default', round(d['ms_per_step'],4), '%.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), 'cpu', d.get('cpu_baseline'))
print({k:v['ms'] for k,v in d['roofline']['stages'].items()})
for a in d.get('also',[]): print(a['workload'], round(a['ms_per_step'],4), '%.4e'%a['value'], a['roofline']['kernel'], round(a['roofline']['frac'],3), 'whole', round(a['whole_step_frac'],3))
PY
