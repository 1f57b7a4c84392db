# GPU session script of the moment (development aid; edited per session) — every command under its own timeout
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_9.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2_bench_1gpu_h.log 2>&1
OON_MORTON_MIN=4096 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2_bench_1gpu_h_morton4096.log 2>&1
tail -n 12 gpurun_out/r2_pytest_9.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_1gpu_h.log","gpurun_out/r2_bench_1gpu_h_morton4096.log"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "value %.4g ms %.4f e2e %.3f ms frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]), d["kernel_ms_per_step"], d["roofline_integrate"]["ms_per_launch"])
    for k,v in d.get("configs",{}).items(): print("   %-32s ms %.4f value %.4g removed %d"%(k,v["ms_per_step"],v["value"],v["removed"]))
PY
