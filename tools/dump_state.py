#!/usr/bin/env python3
"""Development aid: run a synthetic config for some frames on the GPU and dump the body table (for offline analysis)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from out_of_nothing_b200 import synth
from out_of_nothing_b200.world import World
name, n, steps, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
bodies, cfg = synth.make(name, n=n)
w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
w.step(cfg["dt"], steps)
b = w.download()
np.savez_compressed(out, p=b["p"], r=b["r"], v=b["v"])
print("saved", out, b.shape)
