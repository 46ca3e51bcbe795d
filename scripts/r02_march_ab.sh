# A/B timing on one B200: (optionally) the GPU test-suite with the default build, then the default workload once per
# variant.  A variant is NAME or NAME:ENV=VAL[,ENV=VAL...]; a NAME ending in .so is also taken from macos-cfd_b200/lib/.
# usage: [WL=workload] bash scripts/r02_march_ab.sh TAG [notests] variant...
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-ab}; shift
if [ "$1" = "notests" ]; then shift; else
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
fi
for spec in "$@"; do
  name=${spec%%:*}; envs=""
  if [ "$spec" != "$name" ]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  case $name in *.so) envs="$envs CFD_B200_LIB=$PWD/macos-cfd_b200/lib/$name";; esac
  env $envs python bench.py ${WL:+--workload $WL} --steps 12 --warmup 3 --no-also --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${TAG}_$name${WL:+_$WL}.json 2> gpurun_out/${TAG}_$name${WL:+_$WL}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_$name${WL:+_$WL}.json').read().strip().splitlines()[-1])
print('$name', '${WL}', round(d['ms_per_step'],4), d['roofline']['kernel'], round(d['roofline']['ms_per_launch'],4), {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0.05})
PY
done
