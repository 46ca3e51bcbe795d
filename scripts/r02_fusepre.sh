cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "whole_runs or variants or golden or cpp or 1024 or nonfinite or mg_8192 or stage_kernels or shear_pcg" 2>&1 | tail -3
for cfg in "0 libcfd_b200.so" "1 libcfd_b200.so" "1 libcfd_b200_div3.so"; do set -- $cfg
  CFD_B200_FUSE_PRE=$1 CFD_B200_LIB=$PWD/macos-cfd_b200/lib/$2 python bench.py --steps 18 --warmup 3 --no-also --no-e2e --no-cpu-baseline > gpurun_out/r02n_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02n_$1_$2.json").read().strip().splitlines()[-1])
print("fuse_pre=$1 $2", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items() if v["ms"]>0.02})
PY
done
