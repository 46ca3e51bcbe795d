set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
for pdl in 1 0; do
 for wl in jet_mg_1024 jet_mg_8192 lid_pcg_512x256 shear_pcg_4096; do
  st=16; [ $wl = shear_pcg_4096 ] && st=4
  CFD_B200_PDL=$pdl timeout 300 python bench.py --steps $st --warmup 3 --no-e2e --no-cpu-baseline --workload $wl > gpurun_out/pdl${pdl}_$wl.json 2> gpurun_out/pdl${pdl}_$wl.err
  python -c "
import json
d=json.loads(open('gpurun_out/pdl${pdl}_$wl.json').read().strip().splitlines()[-1]); print('PDL',$pdl,'$wl', round(d['ms_per_step'],4), '%.3e'%d['value'], d['gpu_launches'], d['last_step'])"
 done
done
