cd /root/repo; mkdir -p gpurun_out
for pdl in 7 1 3 5 0; do
CFD_B200_PDL=$pdl python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --workload jet_mg_8192 > gpurun_out/rs_pdl$pdl.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/rs_pdl$pdl.json').read().strip().splitlines()[-1]); print('PDL',$pdl, round(d['ms_per_step'],4), d['config']['probe_ms_per_step_from_reset'])
PY
done
