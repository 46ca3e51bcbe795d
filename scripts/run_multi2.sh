cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "jet-mg-1-1-0 or shear-pcg-1-1-0 or lid-pcg-1-1-0 or lid-mg-0-1-0 or jet-mg-1-1-33 or shear-pcg-1-0-0 or jet-mg-1-0-0" > gpurun_out/pytest_multi2.log 2>&1
tail -5 gpurun_out/pytest_multi2.log
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { # name p2p workload steps
  CFD_B200_P2P=$2 timeout 200 $TR bench.py --gpus $N --steps $4 --warmup 3 --no-e2e --no-cpu-baseline --workload $3 > gpurun_out/m${N}_$1.json 2> gpurun_out/m${N}_$1.err
  python - <<PY
import json
t=open('gpurun_out/m${N}_$1.json').read().strip()
if not t: print('$1 EMPTY'); print(open('gpurun_out/m${N}_$1.err').read()[-1500:])
else:
  d=json.loads(t.splitlines()[-1]); print('$1', round(d['ms_per_step'],4), '%.4e'%d['value'], d['config'].get('comm')[:12], d['last_step'])
PY
}
run jet_mg_p2p 1 jet_mg_8192 16
run jet_pcg_p2p 1 jet_pcg_8192 3
run shear_pcg_p2p 1 shear_pcg_4096 4
run shear_pcg_nccl 0 shear_pcg_4096 4
