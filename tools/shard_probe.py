#!/usr/bin/env python3
"""Development aid: what does the interact kernel of a 1/f target shard cost on one GPU, for several launch geometries?

Times the launch one GPU of an f-way sharded world would run (oon_debug_time_interact) on the honest state of the C4
world: every frame advances normally, the partial launches are timed on the side and their sums discarded.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from out_of_nothing_b200 import synth
from out_of_nothing_b200.world import World

n = int(os.environ.get("SWEEP_N", "100000"))
bodies, cfg = synth.make("c4_cloud_100k", n=n)
dt = cfg["dt"]
w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
w.step(dt, 150); w.synchronize()
state = w.download()
w.close()
fracs = [int(x) for x in os.environ.get("PROBE_FRAC", "1,2,4,8").split(",")]
frames = int(os.environ.get("PROBE_FRAMES", "12"))
for group in (int(x) for x in os.environ.get("PROBE_GROUP", "0").split(",")):
    for gpc in (int(x) for x in os.environ.get("PROBE_GPC", "0").split(",")):
        for tail, small in ((int(x), int(y)) for x in os.environ.get("PROBE_TAIL", "6").split(",") for y in os.environ.get("PROBE_SMALL", "-1").split(",")):
            os.environ.pop("OON_GROUP_TILES", None)
            os.environ.pop("OON_SMALL_TILES", None)
            if small >= 0: os.environ["OON_SMALL_TILES"] = str(small)
            if group: os.environ["OON_GROUP_TILES"] = str(group)
            os.environ["OON_DEBUG_GPC"] = str(gpc)
            os.environ["OON_DEBUG_TAIL_X2"] = str(tail)
            w = World(0); w.set_props(**cfg["props"]); w.upload(state)
            w.step(dt, 3)
            acc = {f: [] for f in fracs}
            for k in range(frames):
                w.update_intrinsic(dt)
                for f in fracs:
                    acc[f].append(w.debug_time_interact(dt, f, k % f))
                w.update_pairwise(dt); w.update_global(dt)
            print("group_tiles %d gpc %2d tail_x2 %d small %3d: " % (group, gpc, tail, small) +
                  "  ".join("1/%d %.4f (x%d = %.3f)" % (f, sum(v) / len(v), f, f * sum(v) / len(v)) for f, v in acc.items()), flush=True)
            w.close()
