cd /root/repo; mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --workload jet_mg_1024"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ll1024.csv $CMD > gpurun_out/ncu_l.log 2>&1
python profiles/launch_summary.py gpurun_out/ll1024.csv
