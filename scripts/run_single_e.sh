cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_viz.py tests/test_gpu_cpp_driver.py tests/test_abi.py -m gpu -x -q > gpurun_out/pytest_viz.log 2>&1; tail -25 gpurun_out/pytest_viz.log
