# GPU session script of the moment (development aid; edited per session) — every command under its own timeout
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_7.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2_bench_1gpu_f.log 2>&1
timeout 120 python tools/cta_trace.py 8 6 > gpurun_out/r2_cta_trace_8.log 2>&1
tail -n 5 gpurun_out/r2_pytest_7.log
cat gpurun_out/r2_cta_trace_8.log
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_1gpu_f.log") if l.startswith("{")][-1])
print("value %.4g ms %.4f e2e %.3f ms frac %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["frac"]), d["kernel_ms_per_step"], d["roofline_integrate"]["ms_per_launch"])
for k,v in d["configs"].items(): print("   %-32s ms %.4f value %.4g"%(k,v["ms_per_step"],v["value"]))
PY
for kind in hyperbolic realistic repulsion exact; do timeout 100 python tools/profile_subject.py $kind > gpurun_out/r2_plain_$kind.log 2>&1 || exit 1; done
cat gpurun_out/r2_plain_*.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 160 --csv --log-file gpurun_out/r2_launches_c4.csv python tools/profile_subject.py hyperbolic > gpurun_out/r2_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_interact_fast -s 60 -c 1 -o gpurun_out/r2_prof_interact_hyperbolic python tools/profile_subject.py hyperbolic > gpurun_out/r2_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_interact_fast -s 60 -c 1 -o gpurun_out/r2_prof_interact_realistic python tools/profile_subject.py realistic > gpurun_out/r2_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_interact_fast -s 60 -c 1 -o gpurun_out/r2_prof_interact_repulsion python tools/profile_subject.py repulsion > gpurun_out/r2_ncu_c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 60 -c 1 -o gpurun_out/r2_prof_integrate python tools/profile_subject.py hyperbolic > gpurun_out/r2_ncu_d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_intrinsic_pack -s 60 -c 1 -o gpurun_out/r2_prof_pack python tools/profile_subject.py hyperbolic > gpurun_out/r2_ncu_e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_interact_exact -s 3 -c 1 -o gpurun_out/r2_prof_interact_exact python tools/profile_subject.py exact 6 > gpurun_out/r2_ncu_f.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
