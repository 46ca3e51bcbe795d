cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vcycle.py tests/test_gpu_parity.py -m gpu -x -q -k "vcycle or mg_bit_exact or streaming or full_size" > gpurun_out/pytest_vc.log 2>&1; tail -15 gpurun_out/pytest_vc.log
python scripts/pcie_probe.py 2>&1 | tail -4
