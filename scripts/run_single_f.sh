cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for wl in jet_mg_1024 jet_mg_8192 lid_pcg_512x256; do
  timeout 300 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --workload $wl > gpurun_out/bc_$wl.json 2> gpurun_out/bc_$wl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bc_$wl.json').read().strip().splitlines()[-1]); print('$wl', round(d['ms_per_step'],4), '%.3e'%d['value'], d['gpu_launches'], d['last_step'])
PY
done
