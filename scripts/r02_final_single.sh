# End of round 2 on one B200: the GPU test-suite, smoke(), the default bench line with all its extras, the reference arm,
# the ncu launch list of the default workload and one --set full capture of each hot kernel.  TAG names the output files.
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-r02d}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
( time python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err ) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_default.json').read().strip().splitlines()[-1])
print('default', round(d['ms_per_step'],4), '%.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step']['frac'],3), d['gpu_launches'], 'cpu', d['cpu_baseline']['value'])
print({k:v['ms'] for k,v in d['roofline']['stages'].items()})
for a in d.get('also',[]): print(a['workload'], round(a['ms_per_step'],4), '%.4e'%a['value'], a['roofline']['kernel'][:30], round(a['roofline']['frac'],3))
print(d.get('pressure_solve_times'))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
python profiles/launch_summary.py gpurun_out/${TAG}_launches.csv | head -16
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_rhs_march2|k_project_div|k_rbgs_stream|k_div_fix" -s 12 -c 7 -o gpurun_out/prof_${TAG}_hot $CMD > gpurun_out/ncu_hot.log 2>&1
tail -2 gpurun_out/ncu_hot.log
