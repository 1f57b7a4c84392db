# GPU session script of the moment (development aid; edited per session)
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_2.log 2>&1
SWEEP_WARM=150 python tools/perf_sweep.py > gpurun_out/r2_sweep_t2.log 2>&1
python bench.py > gpurun_out/r2_bench_1gpu.log 2>&1
python bench.py --steps 20 --warmup 5 --skip-extras > gpurun_out/r2_bench_1gpu_driver_style.log 2>&1
python bench.py --steps 20 --warmup 5 --skip-extras --settle 0 --skip-cpu-baseline > gpurun_out/r2_bench_1gpu_nosettle.log 2>&1
tail -n 3 gpurun_out/r2_pytest_2.log
cat gpurun_out/r2_sweep_t2.log
tail -c 1500 gpurun_out/r2_bench_1gpu.log
