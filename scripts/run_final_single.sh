cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/r01c_bench_default.json 2> gpurun_out/r01c_bench_default.err; tail -c 300 gpurun_out/r01c_bench_default.json; echo
for wl in jet_mg_1024 lid_pcg_512x256 shear_pcg_4096; do
  st=16; case $wl in shear*) st=4;; esac
  python bench.py --workload $wl --steps $st --no-cpu-baseline > gpurun_out/r01c_bench_$wl.json 2> gpurun_out/r01c_bench_$wl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r01c_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', round(d['ms_per_step'],4), '%.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3))
PY
done
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
python profiles/launch_summary.py gpurun_out/r01c_launches.csv | head -8
