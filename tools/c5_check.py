#!/usr/bin/env python3
"""Development aid: BASELINE config C5 (1 M bodies) on one GPU — ms/frame over 3 flushed frames and the interact launch a
1/8 shard of it would run (what one GPU of an 8-GPU job does), without needing 8 GPUs."""
import sys, time
sys.path.insert(0, '.')
from out_of_nothing_b200 import synth
from out_of_nothing_b200.world import World
bodies, cfg = synth.make("c5_void_1m")
w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
w.step(cfg["dt"], 1); w.synchronize()
ms = w.bench_frames(cfg["dt"], 3, 256 << 20) / 3
n = bodies.shape[0]
print("c5 1 GPU: %.2f ms/frame, %.3e interactions/s" % (ms, n * (n - 1) / (ms * 1e-3)))
w.update_intrinsic(cfg["dt"])
print("1/8 shard interact launch: %.3f ms" % w.debug_time_interact(cfg["dt"], 8, 3))
w.close()
