#!/usr/bin/env python3
"""Quick on-GPU sanity run (development aid): KAT-1 through the C ABI vs the oracle, exact and fast modes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from out_of_nothing_b200 import snapshot as S, synth
from out_of_nothing_b200.world import World
from out_of_nothing_b200 import capi
import pyoracle as O

def oracle_for(s, bodies, double=False, **kw):
    w = O.OracleWorld(double=double)
    p = dict(friction=s.friction, gravity_mode=s.gravity_mode, gravity=s.gravity, interact_n2n=True); p.update(kw)
    w.set_props(**p); w.set_bodies(bodies); return w

def relerr(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    rms = np.sqrt((b ** 2).sum(axis=1).mean())
    return np.sqrt(((a - b) ** 2).sum(axis=1)) / rms

s = S.load(os.path.join(ROOT, "tests/golden/kat1_start.state"))
e = S.load(os.path.join(ROOT, "tests/golden/kat1_end.state"))
print("fp32 peak TFLOP/s", capi.measure_fp32_peak(0))
for mode in (capi.MATH_EXACT, capi.MATH_FAST):
    w = World(0)
    w.set_props(friction=s.friction, gravity_mode=s.gravity_mode, gravity=s.gravity, interact_n2n=1)
    w.math_mode = mode
    w.debug_capture_pairs(1 << 20)
    w.upload(s.bodies)
    o = oracle_for(s, s.bodies); o.trace(True)
    t0 = time.time(); w.step(0.033, 20); w.synchronize(); t1 = time.time()
    o.step(0.033, 20)
    g = w.download(); ob = o.bodies(); tr = o.trace_fetch()
    ev = w.drain_events()
    gc = w.debug_fetch_pairs(0); gt = w.debug_fetch_pairs(1)
    print("mode", "EXACT" if mode else "FAST", "20 steps %.1f ms" % ((t1 - t0) * 1e3), ev)
    print("  T exact vs oracle", int((g["T"] == ob["T"]).sum()), "/", len(g), " vs golden", int((g["T"] == e.bodies["T"]).sum()))
    print("  p exact", int((g["p"] == ob["p"]).sum()), "v exact", int((g["v"] == ob["v"]).sum()), "of", 2 * len(g))
    print("  p relerr max %.3e med %.3e ; v relerr max %.3e med %.3e" % (relerr(g["p"], ob["p"]).max(), np.median(relerr(g["p"], ob["p"])), relerr(g["v"], ob["v"]).max(), np.median(relerr(g["v"], ob["v"]))))
    print("  oracle coll", len(tr["collisions"]), "touch", len(tr["touches"]), "| gpu coll", len(gc), "touch", len(gt))
    print("  collision multiset equal:", sorted(map(tuple, tr["collisions"].tolist())) == sorted(map(tuple, gc.tolist())), " touch equal:", sorted(map(tuple, tr["touches"].tolist())) == sorted(map(tuple, gt.tolist())))
    w.close()

# quick throughput
for n in (10703, 100000):
    bodies, cfg = synth.make("c4_cloud_100k", n=n)
    w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
    w.step(0.033, 3); w.synchronize()
    k = 20 if n <= 20000 else 5
    w.timer_start(); w.step(0.033, k); ms = w.timer_stop()
    print("N=%d fast: %.3f ms/step  %.3e interactions/s" % (n, ms / k, n * (n - 1) * k / (ms * 1e-3)), w.drain_events())
    prof, launches = w.step_profiled(0.033, 3)
    print("   profiled per step (ms):", {a: b / 3 for a, b in prof.items()}, launches)
    w.close()
