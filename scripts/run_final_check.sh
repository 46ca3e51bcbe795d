cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/r01d_bench_default.json 2> gpurun_out/r01d_bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r01d_bench_default.json').read().strip().splitlines()[-1]); print('default', round(d['ms_per_step'],4), '%.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['gpu_launches'], d['cpu_baseline']['value'])
PY
