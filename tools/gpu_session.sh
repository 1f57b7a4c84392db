# What one gpurun call of this repo usually runs (every command under its own timeout; logs into gpurun_out/).
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -n 2 gpurun_out/pytest_gpu.log
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"; head -c 400 gpurun_out/bench.json
