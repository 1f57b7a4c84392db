# GPU session script of the moment (development aid; edited per session) — 2 GPUs, every command under its own timeout
timeout 400 python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r2_pytest_group_2gpu.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q -k "not multi and not group" > gpurun_out/r2_pytest_5.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2_bench_1gpu_d.log 2>&1
timeout 120 out_of_nothing_b200/host/oon_headless --session=tests/golden/kat1_start.state --interact --friction=0.01 --fixed-dt=0.033 --loop-cap=20 --exit-on-finish --session-save-as=gpurun_out/kat1_end_2gpu.state --no-save-compressed --math=exact --gpus=2 --report > gpurun_out/r2_headless_2gpu.log 2>&1
timeout 120 out_of_nothing_b200/host/oon_headless --session=tests/golden/kat1_start.state --interact --friction=0.01 --fixed-dt=0.033 --loop-cap=20 --exit-on-finish --session-save-as=gpurun_out/kat1_end_1gpu.state --no-save-compressed --math=exact --gpus=1 --report >> gpurun_out/r2_headless_2gpu.log 2>&1
cmp gpurun_out/kat1_end_2gpu.state gpurun_out/kat1_end_1gpu.state >> gpurun_out/r2_headless_2gpu.log 2>&1 && echo "headless --gpus=2 END.state == --gpus=1 END.state (byte for byte)" >> gpurun_out/r2_headless_2gpu.log
tail -n 25 gpurun_out/r2_pytest_group_2gpu.log
tail -n 5 gpurun_out/r2_pytest_5.log
cat gpurun_out/r2_headless_2gpu.log
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_1gpu_d.log") if l.startswith("{")][-1])
print("value %.4g ms %.4f e2e %.3f ms frac %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["frac"]), d["kernel_ms_per_step"], d["roofline_integrate"]["ms_per_launch"])
for k,v in d["configs"].items(): print("   %-32s ms %.4f value %.4g"%(k,v["ms_per_step"],v["value"]))
PY
