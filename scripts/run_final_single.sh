cd /root/repo; mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/r01b_bench_default.json 2> gpurun_out/r01b_bench_default.err; tail -c 600 gpurun_out/r01b_bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01b_bench_reference.json 2>/dev/null; cat gpurun_out/r01b_bench_reference.json
for wl in jet_pcg_8192 lid_mg_8192 shear_pcg_4096 jet_mg_1024 lid_pcg_512x256; do
  st=16; case $wl in *pcg_8192) st=3;; shear*) st=4;; esac
  python bench.py --workload $wl --steps $st --no-cpu-baseline > gpurun_out/r01b_bench_$wl.json 2> gpurun_out/r01b_bench_$wl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r01b_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', round(d['ms_per_step'],4), '%.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['last_step'])
PY
done
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_rbgs_stream -s 4 -c 1 -o gpurun_out/prof_r01b_rbgs_stream $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
