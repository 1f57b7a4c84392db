# GPU session script of the moment (development aid; edited per session) — final 8-GPU bench lines (hot loop unrolled 8x)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2_final_bench_8gpu.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --skip-extras > gpurun_out/r2_final_bench_4gpu.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --skip-extras > gpurun_out/r2_final_bench_2gpu.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2_final_bench_8gpu.log","gpurun_out/r2_final_bench_4gpu.log","gpurun_out/r2_final_bench_2gpu.log"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "value %.4g ms %.4f e2e %.3f ms frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]), d["kernel_ms_per_step"], {k:v for k,v in d.get("parity_check",{}).items() if k!="what"})
    if "c5" in d: print({k:v for k,v in d["c5"].items() if k!="one_gpu"}, d["c5"]["one_gpu"]["ms_per_step"])
PY
