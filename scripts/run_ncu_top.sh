cd /root/repo; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_rbgs_stream -s 4 -c 1 -o gpurun_out/prof_r01d_rbgs_stream $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
