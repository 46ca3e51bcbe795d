cd /root/repo; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/ll.csv $CMD > gpurun_out/ncu_l.log 2>&1
python profiles/launch_summary.py gpurun_out/ll.csv
