# GPU session script of the moment (development aid; edited per session): bench lines of the other BASELINE configs
for wl in c1_kat1 c2_swarm_10k c3_dense_13k; do
  for m in fast exact; do
    timeout 200 python bench.py --workload $wl --math $m --steps 100 --warmup 10 --skip-extras > gpurun_out/r2_bench_${wl}_${m}.log 2>&1
    python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_${wl}_${m}.log") if l.startswith("{")][-1])
print("${wl} ${m}: value %.4g ms %.4f e2e %.3f ms frac %.3f launches %d n_end %d"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"],d["config"]["n_bodies_end"]), d["kernel_ms_per_step"])
PY
  done
done
