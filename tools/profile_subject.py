#!/usr/bin/env python3
"""Subject of the ncu captures under profiles/: `frames` frames of the C4 world (100 000 bodies) through the C ABI.
usage: tools/profile_subject.py [hyperbolic|realistic|repulsion|exact] [frames=70]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from out_of_nothing_b200 import capi, synth  # noqa: E402
from out_of_nothing_b200.world import World  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "hyperbolic"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 70
bodies, cfg = synth.make("c4_cloud_100k", n=20000 if kind == "exact" else None)
props = dict(cfg["props"])
if kind == "realistic":
    props["gravity_mode"] = capi.GRAVITY_REALISTIC
    props["gravity"] = 6.6743e-11 * 1e9          # the inverse-square law needs a stronger constant to move these bodies at all
if kind == "repulsion":
    props["repulsion_stiffness"] = 1e-9
w = World(0)
w.set_props(**props)
if kind == "exact":
    w.math_mode = capi.MATH_EXACT
w.upload(bodies)
w.step(cfg["dt"], frames)
w.synchronize()
ev = w.drain_events()
print("profile subject %s: %d frames, %d bodies, checked fraction %.3f" % (kind, frames, w.entity_count(), ev["warp_tiles_checked"] / max(1, ev["warp_tiles"])))
w.close()
