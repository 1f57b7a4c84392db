mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 240 --csv --log-file gpurun_out/r2_launches_bench_c4.csv python bench.py --steps 20 --warmup 5 > gpurun_out/ncu_bench.log 2>&1; echo "rc $?"
wc -l gpurun_out/r2_launches_bench_c4.csv; tail -n 2 gpurun_out/ncu_bench.log | head -c 400
