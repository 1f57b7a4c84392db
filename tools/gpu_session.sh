# GPU session script of the moment (development aid; edited per session) — final 8-GPU record of round 2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2_final_bench_8gpu.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 > gpurun_out/r2_final_bench_4gpu.log 2>&1
timeout 200 python -m pytest tests/test_gpu_group.py -m gpu -q > gpurun_out/r2_final_pytest_group_8gpu.log 2>&1
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "kat1-0 or boundary_runs or edits or cloud50k-0" > gpurun_out/r2_final_pytest_multi_8gpu.log 2>&1
H=out_of_nothing_b200/host/oon_headless
timeout 120 $H --session=tests/golden/kat1_start.state --interact --friction=0.01 --fixed-dt=0.033 --loop-cap=20 --exit-on-finish --session-save-as=gpurun_out/kat1_end_8gpu.state --no-save-compressed --math=exact --gpus=8 --report > gpurun_out/r2_final_headless_8gpu.log 2>&1
timeout 120 $H --session=tests/golden/kat1_start.state --interact --friction=0.01 --fixed-dt=0.033 --loop-cap=20 --exit-on-finish --session-save-as=gpurun_out/kat1_end_1gpu.state --no-save-compressed --math=exact --gpus=1 --report >> gpurun_out/r2_final_headless_8gpu.log 2>&1
cmp gpurun_out/kat1_end_8gpu.state gpurun_out/kat1_end_1gpu.state >> gpurun_out/r2_final_headless_8gpu.log 2>&1 && echo "headless --gpus=8 END.state == --gpus=1 END.state (byte for byte)" >> gpurun_out/r2_final_headless_8gpu.log
timeout 120 $H --bodies=99999 --interact --friction=0 --fixed-dt=0.033 --loop-cap=400 --gpus=8 --report --no-player-antigravity >> gpurun_out/r2_final_headless_8gpu.log 2>&1
timeout 120 $H --bodies=99999 --interact --friction=0 --fixed-dt=0.033 --loop-cap=400 --gpus=1 --report --no-player-antigravity >> gpurun_out/r2_final_headless_8gpu.log 2>&1
tail -n 5 gpurun_out/r2_final_pytest_group_8gpu.log
tail -n 5 gpurun_out/r2_final_pytest_multi_8gpu.log
cat gpurun_out/r2_final_headless_8gpu.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_final_bench_8gpu.log","gpurun_out/r2_final_bench_4gpu.log"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "value %.4g ms %.4f e2e %.3f ms frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]), d["kernel_ms_per_step"], {k:v for k,v in d["parity_check"].items() if k!="what"})
    if "c5" in d: print({k:v for k,v in d["c5"].items() if k!="one_gpu"}, d["c5"]["one_gpu"]["ms_per_step"])
PY
