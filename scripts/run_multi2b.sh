cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "jet-mg-1-1-33 or lid-mg-1-1-17 or shear-pcg-1-1-0" > gpurun_out/pytest_multi2b.log 2>&1
tail -3 gpurun_out/pytest_multi2b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/m2b_jet_mg.json 2> gpurun_out/m2b.err
python - <<PY
import json
d=json.loads(open('gpurun_out/m2b_jet_mg.json').read().strip().splitlines()[-1]); print('N=2 jet mg', round(d['ms_per_step'],4), '%.4e'%d['value'], d['config'].get('comm')[:12], d['last_step'])
PY
