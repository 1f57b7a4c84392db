"""The oracle pinned on the reference's own golden vectors.

KAT-1 = test/regression/_baseline-2024-09-18/1000_bodies-START.state -> test/regression/tc-smoke/REFERENCE-END.state,
run with tc-smoke.cmd's settings (--interact --friction=0.01 --fixed-dt=0.033 --loop-cap=20).
"""
import numpy as np
import pytest

import pyoracle
from helpers import TC_SMOKE, load_golden, make_oracle, pair_list, props_of, rel_err
from out_of_nothing_b200 import synth


def run_kat1(kat1, double=False, trace=True, **over):
    start, _ = kat1
    o = make_oracle(start.bodies, props_of(start, friction=TC_SMOKE["friction"], interact_n2n=True, **over), double=double, trace=trace)
    o.step(TC_SMOKE["dt"], TC_SMOKE["steps"])
    return o


def test_kat1_float_oracle_reproduces_every_decision(kat1):
    start, end = kat1
    o = run_kat1(kat1)
    got, E = o.bodies(), end.bodies
    tr = o.trace_fetch()
    assert o.n == 1010
    # T bit-exact on every body: all collision/touch decisions of all 20 frames + the cooling law agree
    assert (got["T"] == E["T"]).all()
    assert len(tr["collisions"]) == 3590 and len(tr["touches"]) == 322
    # fields the path must not change
    for f in ("r", "mass", "density", "lifetime", "gravity_immunity", "free_color", "thrust"):
        assert got[f].tobytes() == E[f].tobytes(), f
    # positions / velocities: pinned to ~1 ulp per op (the 2024-09 binary's Vector2 rounding is unrecoverable)
    ep, ev = rel_err(got["p"], E["p"]), rel_err(got["v"], E["v"])
    assert (got["p"] == E["p"]).mean() > 0.55
    assert np.median(ev) < 2e-7 and ev.max() < 2e-5
    assert np.median(ep) < 1e-7 and ep.max() < 1e-5


def test_kat1_double_oracle_same_decisions(kat1):
    o = run_kat1(kat1, double=True)
    tr = o.trace_fetch()
    assert (o.bodies()["T"] == kat1[1].bodies["T"]).all()
    assert len(tr["collisions"]) == 3590 and len(tr["touches"]) == 322


def test_half_and_full_matrix_are_bit_identical_without_repulsion(kat1):
    a = run_kat1(kat1, loop_mode=0, trace=False).bodies()
    b = run_kat1(kat1, loop_mode=1, trace=False).bodies()
    assert a["v"].tobytes() == b["v"].tobytes() and a["p"].tobytes() == b["p"].tobytes()
    # ... except touch heat: full matrix visits each pair twice -> +200 per touching partner (World.cpp:241-249)
    assert (b["T"] >= a["T"]).all() and (b["T"] > a["T"]).any()


def test_update_order_is_pinned(kat1):
    """drag before drift, cooling before touch heat: swapping either breaks the golden."""
    start, end = kat1
    o = make_oracle(start.bodies, props_of(start, friction=TC_SMOKE["friction"], interact_n2n=True))
    for _ in range(TC_SMOKE["steps"]):       # phases called one by one == step()
        o.update_intrinsic(TC_SMOKE["dt"]); o.update_pairwise(TC_SMOKE["dt"]); o.update_global(TC_SMOKE["dt"]); o.sweep_expired()
    assert (o.bodies()["T"] == end.bodies["T"]).all()
    ref = run_kat1(kat1, trace=False).bodies()
    assert o.bodies().tobytes() == ref.tobytes()


def test_radius_formula_reproduces_every_stored_radius(kat1):
    start, _ = kat1
    b = start.bodies
    r = np.array([pyoracle.radius_from_mass_and_density(float(m), float(d)) for m, d in zip(b["mass"], b["density"])], np.float32)
    assert (r == b["r"]).all()
    assert (synth.radius_from_mass_and_density(b["mass"], b["density"]) == b["r"]).all()
    o = pyoracle.OracleWorld(); o.set_bodies(b, recalc=True)
    assert (o.bodies()["r"] == b["r"]).all()      # add_body()'s recalc() is idempotent on fixture data


def test_kat3_statistical_anchor():
    """501 bodies x 500 frames is chaotic (920k colliding pair-steps): only counts, header constants and the
    load-time recalc() are comparable.  The START file is a v0.0.1 snapshot whose stored radii predate the
    mass/density convention; the END radii are what add_body()'s recalc() (World.cpp:64-70) made of them —
    a second, independent pin of the radius formula (501 bodies, none shared with KAT-1)."""
    s, e = load_golden("kat3_start.state"), load_golden("kat3_end.state")
    o = pyoracle.OracleWorld()
    o.set_props(**props_of(s, friction=0.01, interact_n2n=True))
    o.set_bodies(s.bodies, recalc=True)
    assert (o.bodies()["r"] == e.bodies["r"]).all() and (s.bodies["r"] != e.bodies["r"]).all()
    o.step(0.033, 500)
    assert o.n == e.n == 501
    assert abs(e.gravity - 6.6743e-11) < 1e-17
    got = o.bodies()
    assert got["mass"].tobytes() == e.bodies["mass"].tobytes() and got["r"].tobytes() == e.bodies["r"].tobytes()
    assert got["density"].tobytes() == e.bodies["density"].tobytes()


def test_player_only_mode_kat2():
    """interact_n2n = false (the reference's default): only body 0 is a source, but it is kicked by everyone (half matrix)."""
    s = load_golden("kat2_start.state")
    assert not s.interact_n2n
    b = s.bodies.copy(); b["gravity_immunity"][0] = 0
    o = make_oracle(b, props_of(s), trace=True)
    before = o.bodies()
    o.update_pairwise(0.033)
    after, tr = o.bodies(), o.trace_fetch()
    assert tr["pairs_visited"] == s.n - 1
    assert (after["v"][0] != before["v"][0]).any()
    o2 = make_oracle(b, props_of(s, loop_mode=1))
    o2.update_pairwise(0.033)
    assert (o2.bodies()["v"][0] == before["v"][0]).all()     # full matrix: the lone source is never a target


def test_lifetime_termination_and_sweep_quirk():
    """update_intrinsic's expiry (World.cpp:131-141) and the caller's erase-without-index-fix-up (OON.cpp:1089-1095)."""
    b = synth.random_cloud(12)
    b["lifetime"][3:9] = 0.05          # six consecutive bodies expire in the first frame (dt 0.1)
    b["lifetime"][10] = 0.25
    o = make_oracle(b, dict(friction=0.0, interact_n2n=True), trace=True)
    r_before = o.bodies()["r"].copy()
    o.update_intrinsic(0.1)
    mid = o.bodies()
    assert (mid["lifetime"][3:9] == 0).all() and np.allclose(mid["r"][3:9], r_before[3:9] * np.float32(0.1))
    o.update_pairwise(0.1); o.update_global(0.1); o.sweep_expired()
    tr = o.trace_fetch()
    # run of 6 expired starting at 3: the 1st, 3rd, 5th are erased; indices at removal time are 3, 4, 5
    assert tr["expired"].tolist() == [3, 4, 5] and o.n == 9
    o.step(0.1, 1)                     # next frame removes two of the remaining three (again alternate)
    assert o.n == 7
    o.step(0.1, 3)
    assert o.n == 5                    # body 10 (lifetime 0.25) and the last straggler are gone too


def test_float_vs_double_divergence_envelope(kat1):
    """decisions identical through 20 frames; positions within 1e-4 rms."""
    f, d = run_kat1(kat1), run_kat1(kat1, double=True)
    assert pair_list(f.trace_fetch()["collisions"]) == pair_list(d.trace_fetch()["collisions"])
    assert pair_list(f.trace_fetch()["touches"]) == pair_list(d.trace_fetch()["touches"])
    e = rel_err(f.bodies()["p"], d.pv64()[:, :2])
    assert e.max() < 1e-4 and np.median(e) < 1e-5


def test_dt_zero_and_negative(kat1):
    start, _ = kat1
    o = make_oracle(start.bodies, props_of(start, interact_n2n=True))
    o.update_intrinsic(0.0)
    assert o.bodies().tobytes() == start.bodies.tobytes()          # World.cpp:108-111
    o.step(-0.033, 1)                                              # time reversal (OON_sfml.cpp:152) must not blow up
    assert np.isfinite(o.bodies()["p"]).all()
