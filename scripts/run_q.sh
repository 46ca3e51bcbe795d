cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mg or streaming or full_size or nonfinite or variants" > gpurun_out/pytest_q.log 2>&1; tail -2 gpurun_out/pytest_q.log
python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/rs_default.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/rs_default.json').read().strip().splitlines()[-1]); print('default', round(d['ms_per_step'],4), d['roofline']['stages']['pressure_solve'], d['config']['probe_ms_per_step_from_reset'])
PY
