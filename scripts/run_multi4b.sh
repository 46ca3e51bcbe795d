cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "jet-mg-1-1-0 or shear-pcg-1-1-0" > gpurun_out/pytest_multi4b.log 2>&1
tail -3 gpurun_out/pytest_multi4b.log
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
CFD_B200_P2P=1 timeout 200 $TR bench.py --gpus $N --steps 12 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/m4b_jet_mg_p2p.json 2> gpurun_out/m4b.err
CFD_B200_P2P=1 timeout 200 $TR bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload jet_pcg_8192 > gpurun_out/m4b_jet_pcg_p2p.json 2>> gpurun_out/m4b.err
python - <<PY
import json
for f in ('m4b_jet_mg_p2p','m4b_jet_pcg_p2p'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), '%.4e'%d['value'], d['config'].get('comm')[:12], d['last_step'])
PY
