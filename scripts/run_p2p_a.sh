set -x
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "jet-mg-1-1 or shear-pcg-1-1 or lid-mg-0-1 or jet-mg-1-0" > gpurun_out/pytest_multi_p2p.log 2>&1
tail -5 gpurun_out/pytest_multi_p2p.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for mode in 1 0; do
  CFD_B200_P2P=$mode timeout 200 $TR bench.py --gpus 2 --steps 16 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/p2p${mode}_n2_jet_mg.json 2> gpurun_out/p2p${mode}_n2_jet_mg.err
  CFD_B200_P2P=$mode timeout 200 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload jet_pcg_8192 > gpurun_out/p2p${mode}_n2_jet_pcg.json 2> gpurun_out/p2p${mode}_n2_jet_pcg.err
  CFD_B200_P2P=$mode timeout 200 $TR bench.py --gpus 2 --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --workload lid_mg_8192_strong > gpurun_out/p2p${mode}_n2_lid_strong.json 2> gpurun_out/p2p${mode}_n2_lid_strong.err
done
for f in gpurun_out/p2p*_n2_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['ms_per_step'], d['value'], d['config'].get('comm'), d['last_step'])"; done
tail -3 gpurun_out/p2p*_n2_*.err
