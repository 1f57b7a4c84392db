# GPU session script of the moment (development aid; edited per session) — 2 GPUs, every command under its own timeout
timeout 400 python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r2_pytest_group_2gpu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "kat1 or boundary_runs or edits or cloud50k-0 or c3_dense-1" > gpurun_out/r2_pytest_multi_2gpu.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2_bench_2gpu.log 2>&1
tail -n 12 gpurun_out/r2_pytest_group_2gpu.log
tail -n 12 gpurun_out/r2_pytest_multi_2gpu.log
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_2gpu.log") if l.startswith("{")][-1])
print("value %.4g ms %.4f e2e %.3f ms frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]), d["kernel_ms_per_step"], d["parity_check"])
PY
