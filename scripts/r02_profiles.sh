# round 2: the ncu evidence of the default workload -- launch list (shares) and one --set full capture of each hot kernel
cd /root/repo; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
python profiles/launch_summary.py gpurun_out/r02_launches.csv | head -14
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_rhs_march|k_project_div|k_divergence|k_rbgs_stream|k_div_fix" -s 24 -c 7 -o gpurun_out/prof_r02_hot $CMD > gpurun_out/ncu_hot.log 2>&1
tail -2 gpurun_out/ncu_hot.log
