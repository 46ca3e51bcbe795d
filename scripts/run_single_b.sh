cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_vcycle.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
for v in 1 2; do
 for wl in jet_mg_8192 jet_mg_1024; do
  timeout 300 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --workload $wl --variant $v > gpurun_out/rs_v${v}_$wl.json 2> gpurun_out/rs_v${v}_$wl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/rs_v${v}_$wl.json').read().strip().splitlines()[-1]); print('variant',$v,'$wl', round(d['ms_per_step'],4), '%.3e'%d['value'], d['gpu_launches'], d['last_step'], d['roofline']['stages']['pressure_solve'])
PY
 done
done
for ry in 64 128 512; do
  CFD_B200_RB_RY=$ry timeout 300 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --workload jet_mg_8192 > gpurun_out/rs_ry${ry}.json 2> gpurun_out/rs_ry${ry}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/rs_ry${ry}.json').read().strip().splitlines()[-1]); print('RY',$ry, round(d['ms_per_step'],4), d['roofline']['stages']['pressure_solve'])
PY
done
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mg_bit_exact and (53 or 421)" > gpurun_out/san_memcheck.log 2>&1; tail -5 gpurun_out/san_memcheck.log
