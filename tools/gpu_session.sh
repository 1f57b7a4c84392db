mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log
timeout 120 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc $?" >> gpurun_out/final_bench.err
tail -3 gpurun_out/final_pytest.log; cat gpurun_out/final_bench.json | head -c 600
