# GPU session script of the moment (development aid; edited per session) — every command under its own timeout
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_11.log 2>&1
tail -n 4 gpurun_out/r2_pytest_11.log
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-extras > gpurun_out/r2_bench_1gpu_j.log 2>&1
OON_HEAVY_FIRST=0 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-extras > gpurun_out/r2_bench_1gpu_j_off.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_1gpu_j.log","gpurun_out/r2_bench_1gpu_j_off.log"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "value %.4g ms %.4f frac %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["frac"]), d["kernel_ms_per_step"])
PY
timeout 120 python tools/cta_trace.py 8 6 > gpurun_out/r2_cta_trace_8_heavy_first.log 2>&1; grep -v "tiles: duration" gpurun_out/r2_cta_trace_8_heavy_first.log
OON_HEAVY_FIRST=0 timeout 120 python tools/cta_trace.py 8 6 2>&1 | tail -2
timeout 120 python tools/cta_trace.py 4 4 2>&1 | tail -1
OON_HEAVY_FIRST=0 timeout 120 python tools/cta_trace.py 4 4 2>&1 | tail -1
timeout 120 python tools/cta_trace.py 2 4 2>&1 | tail -1
OON_HEAVY_FIRST=0 timeout 120 python tools/cta_trace.py 2 4 2>&1 | tail -1
