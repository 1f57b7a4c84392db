"""Shared helpers of the test-suite (test infrastructure; may use the oracle)."""
import os

import numpy as np

import pyoracle
from out_of_nothing_b200 import snapshot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# the reference's own smoke-test settings (test/regression/tc-smoke.cmd:30-43)
TC_SMOKE = dict(dt=0.033, steps=20, friction=0.01)

# ---- stated tolerances (SURVEY.md §8d "Parity gates"), FP32, relative to the rms magnitude over all bodies
TOL_KICK_MEDIAN = 2e-6     # teacher-forced, per step: kick and velocity, median over bodies
TOL_KICK_MAX = 5e-5        # ... maximum over bodies
TOL_POS_MAX = 1e-6         # teacher-forced, per step: position, maximum over bodies
TOL_FREE_POS_MAX = 1e-4    # free-running 20 frames vs the double-precision oracle: position max
TOL_FREE_POS_MEDIAN = 1e-5 # ... median


def load_golden(name):
    return snapshot.load(os.path.join(GOLDEN, name))


def props_of(snap, **over):
    p = dict(friction=snap.friction, gravity_mode=snap.gravity_mode, gravity=snap.gravity,
             interact_n2n=snap.interact_n2n, repulsion_stiffness=0.0, collision_mode=1, loop_mode=0)
    p.update(over)
    return p


def make_oracle(bodies, props, double=False, trace=False, recalc=False):
    """recalc=True: the bodies go through add_body()'s recalc() like World::load does (World_SaveLoad.cpp:156)."""
    o = pyoracle.OracleWorld(double=double)
    o.set_props(**props)
    o.set_bodies(bodies, recalc=recalc)
    if trace:
        o.trace(True)
    return o


def rel_err(a, b):
    """per-body |a-b|_2 / rms_over_bodies(|b|_2)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    rms = np.sqrt((b ** 2).sum(axis=1).mean())
    if rms == 0:
        rms = 1.0
    return np.sqrt(((a - b) ** 2).sum(axis=1)) / rms


def pos_tol(ref_bodies, dt):
    """Position bound for one frame relative to rms(|p|): the stated position tolerance plus what the velocity
    tolerance contributes through the drift p += v*dt (dominant when kicks dwarf the positions' own scale)."""
    p = ref_bodies["p"].astype(np.float64); v = ref_bodies["v"].astype(np.float64)
    rms_p = np.sqrt((p ** 2).sum(axis=1).mean()) or 1.0
    rms_v = np.sqrt((v ** 2).sum(axis=1).mean())
    return TOL_POS_MAX + TOL_KICK_MAX * abs(dt) * rms_v / rms_p


def pair_list(a):
    return sorted(map(tuple, np.asarray(a).reshape(-1, 2).tolist()))


def bodies_equal_except_color(a, b):
    """bitwise equality of every field but `color` (T_to_RGB_and_BV is absent from the reference tree)."""
    for f in a.dtype.names:
        if f in ("color", "_pad"):
            continue
        if a[f].tobytes() != b[f].tobytes():
            return False, f
    return True, None


def kick_f64_sample(bodies, props, dt, targets):
    """Velocity change of one pairwise pass for the given target indices, evaluated in float64 with numpy.
    An independent, directed restatement of World.cpp:271-288,304,358-364 used where the oracle itself
    (O(N^2) on one core) is too slow: full-size worlds, sampled targets.  Half-/full-matrix agree when
    repulsion is off; lifetimes are assumed unlimited."""
    p = bodies["p"].astype(np.float64)
    m = bodies["mass"].astype(np.float64)
    r = bodies["r"].astype(np.float64)
    G = float(np.float32(props["gravity"]))
    out = np.zeros((len(targets), 2))
    for k, i in enumerate(targets):
        if bodies["gravity_immunity"][i]:
            continue
        d = p - p[i]
        dist = np.hypot(d[:, 0], d[:, 1])
        ok = dist > (r + r[i]) if props.get("collision_mode", 1) else np.ones(len(p), bool)
        ok[i] = False
        mode = props.get("gravity_mode", 1)
        g = G / dist[ok] if mode == 1 else (G / dist[ok] ** 2 if mode in (2, 3) else 0.0 * dist[ok])
        out[k] = ((d[ok] / dist[ok, None]) * (g * m[ok])[:, None]).sum(axis=0) * float(np.float32(dt))
    return out
