# one `ncu --set full` capture of one kernel of the default workload.  usage: bash scripts/r02_ncu.sh TAG KERNEL_REGEX [skip]
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-r02}; KRE=${2:-k_rbgs_stream}; SKIP=${3:-4}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
