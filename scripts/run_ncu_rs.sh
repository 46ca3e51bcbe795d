cd /root/repo; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload jet_mg_8192"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_rbgs -s 8 -c 2 -o gpurun_out/prof_r01_rs3 $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
