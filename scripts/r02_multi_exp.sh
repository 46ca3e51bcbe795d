# N-GPU timing experiments with environment knobs.  usage: r02_multi_exp.sh N TAG "ENV1=.. ENV2=.." ["ENV.."] ...
cd /root/repo; mkdir -p gpurun_out
N=$1; TAG=$2; shift; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus $N --steps 20 --warmup 5 --no-also --no-e2e ${WL:+--workload $WL} > gpurun_out/${TAG}_exp$i.json 2> gpurun_out/${TAG}_exp$i.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_exp$i.json').read().strip().splitlines()[-1])
print("$envs", "N=$N", round(d["ms_per_step"],4), d["config"].get("peer_wait_us_per_step_rank0"), "parity_ok", (d.get("parity") or {}).get("ok"), {k:v["ms"] for k,v in d["roofline"]["stages"].items()})
PY
done
