mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_group.py -x -q -m gpu > gpurun_out/final_group_2gpu.log 2>&1; echo "rc $?" >> gpurun_out/final_group_2gpu.log
timeout 100 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "kat1 or edits or boundary_runs" > gpurun_out/final_multi_2gpu.log 2>&1; echo "rc $?" >> gpurun_out/final_multi_2gpu.log
tail -n 3 gpurun_out/final_group_2gpu.log gpurun_out/final_multi_2gpu.log
