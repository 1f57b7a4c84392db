# GPU session script of the moment (development aid; edited per session) — 2 GPUs
timeout 400 python -m pytest tests/test_gpu_group.py -m gpu -q > gpurun_out/r2_final_pytest_group_2gpu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2_final_pytest_multi_2gpu.log 2>&1
tail -n 6 gpurun_out/r2_final_pytest_group_2gpu.log
tail -n 6 gpurun_out/r2_final_pytest_multi_2gpu.log
