cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mg or variants or streaming or nonfinite or full_size" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for wl in jet_mg_8192; do
  timeout 300 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --workload $wl > gpurun_out/rs_$wl.json 2> gpurun_out/rs_$wl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/rs_$wl.json').read().strip().splitlines()[-1]); print('$wl', round(d['ms_per_step'],4), '%.3e'%d['value'], d['gpu_launches'], d['roofline']['stages']['pressure_solve'], d['config']['probe_ms_per_step_from_reset'])
PY
done
