#!/usr/bin/env python3
"""Development aid: steady-state ms/step of the C4 world for each library variant (run each in a fresh process)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
from out_of_nothing_b200 import synth
from out_of_nothing_b200.world import World
n = int(os.environ.get("SWEEP_N", "100000"))
bodies, cfg = synth.make("c4_cloud_100k", n=n)
scale = float(os.environ.get("SWEEP_SCALE", "1"))
if scale != 1:
    bodies["p"] *= scale          # spread the cloud: no collisions, (almost) every tile culled -> speed of the unchecked loop
w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
w.step(cfg["dt"], int(os.environ.get("SWEEP_WARM", "150"))); w.synchronize()
best = 1e9
for rep in range(3):
    w.timer_start(); w.step(cfg["dt"], 32); ms = w.timer_stop() / 32
    best = min(best, ms)
ev = w.drain_events()
prof, _ = w.step_profiled(cfg["dt"], 16)
print("%%.4f ms/step  interact %%.4f  checked %%.3f" %% (best, prof["interact"] / 16, ev["warp_tiles_checked"] / max(1, ev["warp_tiles"])))
''' % ROOT
libs = sorted(glob.glob(os.path.join(ROOT, "out_of_nothing_b200/lib/variants/*.so"))) or [os.path.join(ROOT, "out_of_nothing_b200/lib/liboon_gpu.so")]
for lib in libs:
    env = dict(os.environ, OON_GPU_LIB=lib)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("%-12s %s %s" % (os.path.basename(lib)[:-3], out.stdout.strip(), out.stderr.strip()[-200:]))
