# one --set full capture of each marching kernel of the default workload (ncu replays: never a bench number)
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-r02q}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-parity"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_rhs_march" -s 8 -c 2 -o gpurun_out/prof_${TAG}_march $CMD > gpurun_out/ncu_${TAG}_march.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_march.log
