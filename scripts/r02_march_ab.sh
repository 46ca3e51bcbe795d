# A/B of the marching advection kernel on one B200: the GPU test-suite with the default (k_rhs_march2), then the default
# workload with the first version (CFD_B200_MARCH=1), the default build, and alternative builds given as arguments.
# usage: bash scripts/r02_march_ab.sh TAG [lib.so ...]
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-ab}; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 12 --warmup 3 --no-also --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
print('$name', round(d['ms_per_step'],4), d['roofline']['kernel'], round(d['roofline']['ms_per_launch'],4), {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0.1})
PY
}
run march1 CFD_B200_MARCH=1
run march2 CFD_B200_MARCH=2
for lib in "$@"; do run $lib CFD_B200_LIB=$PWD/macos-cfd_b200/lib/$lib; done
