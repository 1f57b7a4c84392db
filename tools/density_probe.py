#!/usr/bin/env python3
"""Development aid: colliding-pair fraction and step time over a run of a synthetic config."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from out_of_nothing_b200 import synth
from out_of_nothing_b200.world import World
import os
RUNS = (("c4_cloud_100k", 100000, 20, 220), ("c2_swarm_10k", 10703, 50, 300), ("c3_dense_13k", 13503, 50, 200))
if os.environ.get("PROBE_ONLY"): RUNS = tuple(r for r in RUNS if r[0] == os.environ["PROBE_ONLY"])
for name, n, every, total in RUNS:
    bodies, cfg = synth.make(name, n=n)
    w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
    print(name, "N", n)
    done = 0
    while done < total:
        w.timer_start(); w.step(cfg["dt"], every); ms = w.timer_stop(); done += every
        ev = w.drain_events(); nn = w.entity_count()
        pairs = nn * (nn - 1) / 2 * every
        b = w.download()
        print("  step %4d N %6d: %.3f ms/step  checked tiles %.3f  coll frac %.3e touch frac %.3e  |p| med %.2e max %.2e |v| med %.2e" % (
            done, nn, ms / every, ev["warp_tiles_checked"] / max(1, ev["warp_tiles"]), ev["collisions"] / pairs, ev["touches"] / pairs,
            np.median(np.hypot(b["p"][:, 0], b["p"][:, 1])), np.abs(b["p"]).max(), np.median(np.hypot(b["v"][:, 0], b["v"][:, 1]))))
    w.close()
