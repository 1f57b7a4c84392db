# GPU session script of the moment (development aid; edited per session)
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_12.log 2>&1
tail -n 12 gpurun_out/r2_pytest_12.log
python - <<'PY'
# a 100k-body world with a thrusting player's exhaust: frames/s with the rebuilt curve order vs the identity order (OON_MORTON=0)
import os, subprocess, sys
CHILD = r"""
import os, sys
sys.path.insert(0, os.getcwd())
from out_of_nothing_b200 import synth
from out_of_nothing_b200.world import World
bodies, cfg = synth.make("c4_cloud_100k")
bodies = bodies.copy(); bodies["thrust"][0] = [3e32, 0, 1e32, 0]
w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
ecfg = w.emitter_config(particle_lifetime=0.12, create_mass=0, particle_mass_min=1e20, particle_mass_max=3e20, eject_velocity=(0.0, -4e8))
import time
for k in range(30):
    w.emit_particles(0, 5, ecfg, seed=k); w.step(cfg["dt"], 1)
w.synchronize(); t0 = time.perf_counter()
for k in range(30, 90):
    w.emit_particles(0, 5, ecfg, seed=k); w.step(cfg["dt"], 1)
w.synchronize(); dt_ = time.perf_counter() - t0
ev = w.drain_events()
print("%s: %.3f ms/frame, %d bodies, removed %d, checked fraction %.3f" % (os.environ.get("OON_MORTON", "1"), dt_ / 60 * 1e3, w.entity_count(), ev["removed"], ev["warp_tiles_checked"] / max(1, ev["warp_tiles"])))
"""
for m in ("1", "0"):
    out = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, OON_MORTON=m), capture_output=True, text=True)
    print("OON_MORTON=" + out.stdout.strip(), out.stderr[-300:])
PY
