# GPU session script of the moment (development aid; edited per session) — final 1-GPU record (hot loop unrolled 8x)
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_1gpu.log 2>&1
tail -n 4 gpurun_out/r2_final_pytest_1gpu.log
timeout 400 python bench.py > gpurun_out/r2_final_bench_1gpu.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_1gpu_driver_style.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2_final_bench_1gpu.log","gpurun_out/r2_final_bench_1gpu_driver_style.log"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "value %.4g ms %.4f"%(d["value"],d["ms_per_step"]), "e2e", d["e2e"].get("ms_per_step"), d.get("roofline",{}).get("frac"), d.get("kernel_ms_per_step"), d.get("cpu_baseline",{}) and d["cpu_baseline"]["value"], d["first_frames"])
    for k,v in d.get("configs",{}).items(): print("   %-32s ms %.4f value %.4g"%(k,v["ms_per_step"],v["value"]))
PY
timeout 100 python tools/profile_subject.py hyperbolic 40 > gpurun_out/r2_plain_final.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 90 --csv --log-file gpurun_out/r2_final_launches_c4.csv python tools/profile_subject.py hyperbolic 40 > gpurun_out/r2_ncu_launches_final.log 2>&1
mkdir -p /tmp/prof; timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_interact_fast -s 30 -c 1 -o /tmp/prof/k_interact_fast_final python tools/profile_subject.py hyperbolic 40 > gpurun_out/r2_ncu_final.log 2>&1; python tools/ncu_summary.py /tmp/prof/k_interact_fast_final.ncu-rep --top > gpurun_out/r2_k_interact_fast_hyperbolic_ncu_summary.txt 2>&1
head -16 gpurun_out/r2_k_interact_fast_hyperbolic_ncu_summary.txt
