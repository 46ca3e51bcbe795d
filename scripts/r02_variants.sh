# time (and optionally test) the default workload with alternative builds of the library (CFD_B200_LIB).
# usage: r02_variants.sh TAG lib1.so lib2.so ...   (PYTEST_K="expr" also runs that pytest selection with each build)
cd /root/repo; mkdir -p gpurun_out
TAG=$1; shift
for lib in "$@"; do
  export CFD_B200_LIB=$PWD/macos-cfd_b200/lib/$lib
  if [ -n "$PYTEST_K" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$PYTEST_K" 2>&1 | tail -2; fi
  python bench.py --steps 12 --warmup 3 --no-also --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${TAG}_$lib.json 2> gpurun_out/${TAG}_$lib.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_$lib.json').read().strip().splitlines()[-1])
print('$lib', round(d['ms_per_step'],4), d['roofline']['kernel'], round(d['roofline']['ms_per_launch'],4), {k:v['ms'] for k,v in d['roofline']['stages'].items() if v['ms']>0.1})
PY
done
