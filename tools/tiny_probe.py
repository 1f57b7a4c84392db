#!/usr/bin/env python3
"""Development aid: step a small synthetic world a few frames (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from out_of_nothing_b200 import synth
from out_of_nothing_b200.world import World
n = int(sys.argv[1]); steps = int(sys.argv[2])
bodies, cfg = synth.make("c4_cloud_100k", n=n)
w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
w.step(cfg["dt"], steps); w.synchronize()
print(w.drain_events())
