#!/usr/bin/env python3
"""Development aid: what does the interact kernel of a 1/f target shard cost on one GPU, for several launch geometries?

Emulates the per-GPU launch of a sharded world (OON_DEBUG_TARGET_FRACTION launches 1/f of the target blocks) at the
steady-state density of the C4 world.  Timing only: the frames run under the fraction are not physical.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from out_of_nothing_b200 import synth
from out_of_nothing_b200.world import World

n = int(os.environ.get("SWEEP_N", "100000"))
bodies, cfg = synth.make("c4_cloud_100k", n=n)
w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
w.step(cfg["dt"], 150); w.synchronize()
state = w.download()
prof, _ = w.step_profiled(cfg["dt"], 8)
print("full world: interact %.4f ms" % (prof["interact"] / 8))
w.close()

for frac in (int(x) for x in os.environ.get('PROBE_FRAC', '2,4,8').split(',')):
    for group in (int(x) for x in os.environ.get('PROBE_GROUP', '0').split(',')):
        for cpc in (int(x) for x in os.environ.get('PROBE_GPC', '0,1,2,4').split(',')):
            if group: os.environ["OON_GROUP_TILES"] = str(group)
            os.environ["OON_DEBUG_GPC"] = str(cpc)
            os.environ.pop("OON_DEBUG_TARGET_FRACTION", None)
            w = World(0); w.set_props(**cfg["props"]); w.upload(state)
            w.step(cfg["dt"], 2); w.synchronize()
            if frac > 1: os.environ["OON_DEBUG_TARGET_FRACTION"] = str(frac)
            parts = []
            for part in range(frac if os.environ.get("PROBE_PARTS") else 1):
                os.environ["OON_DEBUG_TARGET_PART"] = str(part)
                nf = int(os.environ.get("PROBE_FRAMES", "2"))
                best = 1e9
                for rep in range(3 if nf < 10 else 1):
                    prof, _ = w.step_profiled(cfg["dt"], nf)
                    best = min(best, prof["interact"] / nf)
                parts.append(best)
            print("1/%d shard  group_tiles %d gpc %2d: interact %s ms (ideal %.4f)" % (frac, group, cpc, " ".join("%.4f" % x for x in parts), 3.02 / frac))
            w.close()
