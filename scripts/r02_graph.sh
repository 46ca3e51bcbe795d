cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "whole_runs or variants or async or golden or cpp or 1024 or nonfinite or mg_8192" 2>&1 | tail -3
for g in 0 1; do for wl in jet_mg_1024 jet_mg_8192 lid_mg_8192x1024; do
  CFD_B200_GRAPH=$g python bench.py --workload $wl --steps 24 --warmup 5 --no-also --no-e2e --no-cpu-baseline > gpurun_out/r02m_g${g}_$wl.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02m_g${g}_$wl.json").read().strip().splitlines()[-1])
print("graph=$g $wl", round(d["ms_per_step"],4), d["gpu_launches"])
PY
done; done
