# GPU session script of the moment (development aid; edited per session) — every command under its own timeout
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_3.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2_bench_1gpu_b.log 2>&1
tail -n 15 gpurun_out/r2_pytest_3.log
tail -c 600 gpurun_out/r2_bench_1gpu_b.log
