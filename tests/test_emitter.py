"""Emitter::emit_particles (/root/reference/src/app/model/Emitter.cpp:22-92) — SURVEY.md §8 row f3.

CPU part: the oracle's restatement (rand() replaced by the counter RNG the C ABI documents).
GPU part: the device-side emitter appends bit-identical bodies, and a scene that emits every frame while particles
expire and are swept stays bit-identical to the oracle (EXACT) / keeps every decision (FAST).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402

from helpers import bodies_equal_except_color, load_golden, make_oracle, pair_list, props_of  # noqa: E402
from out_of_nothing_b200 import capi, synth  # noqa: E402


def exhaust_cfg(**over):
    """exhaust_burst's emitter (OON.cpp:884-897): finite lifetime, light particles, mass taken from / given to the emitter."""
    c = capi.EmitterCfg()
    c.v_factor, c.offset_velo_factor = 0.1, 0.2
    c.particle_lifetime, c.create_mass = 1.0, 1
    c.particle_density = 2.0
    c.position_divergence_x = c.position_divergence_y = 1.0
    c.velocity_divergence = 1.0
    c.particle_mass_min, c.particle_mass_max = 2.3e20, 2.9e19      # (the reference's min/max size ratios are swapped too)
    c.eject_velocity_x, c.eject_velocity_y = 0.0, -4e9
    c.eject_offset_x, c.eject_offset_y = 0.0, -1.5e8
    c.color = 0xaaaaaa
    for k, v in over.items():
        setattr(c, k, v)
    return c


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_counter_rng_known_answers():
    """Pins oon_rand (splitmix64 of seed ^ k*odd, top 15 bits): the device kernel and the oracle must agree on these."""
    L = pyoracle.lib()
    assert [L.oracle_rand(7, k) for k in range(8)] == [12773, 27809, 8426, 9163, 22768, 21658, 26960, 21854]
    draws = np.array([L.oracle_rand(20240918, k) for k in range(20000)])
    assert draws.min() >= 0 and draws.max() <= 32767 and abs(draws.mean() - 16383.5) < 200


def test_oracle_emitter_follows_the_reference_expressions():
    s = load_golden("kat1_start.state")
    o = make_oracle(s.bodies, props_of(s, interact_n2n=True))
    b0 = o.bodies()[0].copy()
    c = exhaust_cfg()
    assert o.emit_particles(0, 7, c, None, seed=99) == 7 and o.n == 1017
    new = o.bodies()[1010:]
    assert (new["lifetime"] == np.float32(1.0)).all() and (new["density"] == np.float32(2.0)).all() and (new["T"] == 0).all()
    assert (new["thrust"] == np.float32(2e31)).all() and (new["gravity_immunity"] == 0).all()
    rm = np.float32(32767)
    for i, body in enumerate(new):
        r = [np.float32(pyoracle.lib().oracle_rand(99, 5 * i + j)) for j in range(5)]
        mass = np.float32(c.particle_mass_min) + (np.float32(c.particle_mass_max) - np.float32(c.particle_mass_min)) * r[0] / rm
        assert body["mass"] == mass
        assert body["r"] == np.float32(pyoracle.radius_from_mass_and_density(mass, 2.0))
        pr = np.float32(c.position_divergence_x) * b0["r"]
        vr = b0["r"] * np.float32(c.velocity_divergence)
        # the emitter of KAT-1 is at rest: normalized() of the null vector adds nothing to the offset
        assert body["p"][0] == (r[1] * pr) / rm - pr / np.float32(2) + b0["p"][0] + np.float32(c.eject_offset_x)
        assert body["v"][1] == (r[4] * vr) / rm - vr / np.float32(2) + b0["v"][1] * np.float32(c.v_factor) + np.float32(c.eject_velocity_y)
    assert o.bodies()[0]["mass"] == b0["mass"]                   # create_mass: the emitter keeps its mass


def test_oracle_emitter_depletes_and_skips():
    bodies, cfg = synth.make("c4_cloud_100k", n=20)
    o = make_oracle(bodies, cfg["props"])
    m0 = float(bodies[3]["mass"])
    c = exhaust_cfg(create_mass=0, particle_mass_min=m0 * 0.2, particle_mass_max=m0 * 0.4)
    k = o.emit_particles(3, 10, c, None, seed=5)
    out = o.bodies()
    assert 2 <= k <= 4 and o.n == 20 + k                          # the emitter runs out of mass: the rest is skipped
    given = out[20:]["mass"].astype(np.float64).sum()
    assert abs(float(out[3]["mass"]) + given - m0) <= 1e-6 * m0
    assert out[3]["r"] == np.float32(pyoracle.radius_from_mass_and_density(out[3]["mass"], out[3]["density"]))


# ------------------------------------------------------------------------------------------------ GPU
def gpu_world(bodies, props, mode):
    from out_of_nothing_b200.world import World
    w = World(0)
    w.set_props(**{k: (int(v) if isinstance(v, bool) else v) for k, v in props.items()})
    w.math_mode = mode
    w.debug_capture_pairs(1 << 20)
    w.upload(bodies)
    return w


@pytest.mark.gpu
@pytest.mark.parametrize("create_mass", [1, 0])
@pytest.mark.parametrize("n", [1, 5, 300, 1000])
def test_device_emitter_is_bit_identical_to_the_oracle(n, create_mass):
    bodies, cfg = synth.make("c4_cloud_100k", n=700)
    bodies = bodies.copy()
    bodies["v"][5] = (3e8, -4e8)                                  # a moving emitter: normalized() matters
    rng = np.random.default_rng(n)
    nozzles = rng.uniform(-1, 1, (n, 2)).astype(np.float32) if n == 300 else None
    m5 = float(bodies[5]["mass"])
    c = exhaust_cfg(create_mass=create_mass, particle_mass_min=m5 * 1e-3, particle_mass_max=m5 * (3e-3 if n < 1000 else 4e-3))
    o = make_oracle(bodies, cfg["props"])
    w = gpu_world(bodies, cfg["props"], capi.MATH_EXACT)
    for burst, emitter in enumerate((5, 0, 5)):
        ko = o.emit_particles(emitter, n, c, nozzles, seed=1000 + burst)
        kg = w.emit_particles(emitter, n, c, nozzles, seed=1000 + burst)
        assert ko == kg
    assert w.entity_count() == o.n
    ok, field = bodies_equal_except_color(w.download(), o.bodies())
    assert ok, field
    if not create_mass and n == 1000:
        assert o.n < 700 + 3000                                   # some particles were refused: the emitter ran dry
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [capi.MATH_EXACT, capi.MATH_FAST])
def test_thrusting_player_scene_with_expiring_exhaust(mode):
    """exhaust_burst every frame (OON.cpp:866-950) while the particles of earlier frames expire and are swept."""
    s = load_golden("kat1_start.state")
    bodies = s.bodies.copy()
    bodies["thrust"][0, 0] = 3e36                                 # the player is firing its 'up' thruster
    props = props_of(s, friction=0.01, interact_n2n=True)
    c = exhaust_cfg(particle_lifetime=0.12)                        # ~4 frames at dt = 0.033
    o = make_oracle(bodies, props, trace=True)
    w = gpu_world(bodies, props, mode)
    dt = 0.033
    for frame in range(30):
        assert o.emit_particles(0, 5, c, None, seed=frame) == w.emit_particles(0, 5, c, None, seed=frame)
        if mode == capi.MATH_FAST:                                # teacher-forced: same input state every frame
            w.upload(o.bodies())
        o.trace_clear()
        o.step(dt, 1)
        w.step(dt, 1)
        tr = o.trace_fetch()
        assert w.entity_count() == o.n
        assert pair_list(w.debug_fetch_pairs(0)) == pair_list(tr["collisions"])
        assert pair_list(w.debug_fetch_pairs(1)) == pair_list(tr["touches"])
        got, ref = w.download(), o.bodies()
        if mode == capi.MATH_EXACT:
            ok, field = bodies_equal_except_color(got, ref)
            assert ok, (frame, field)
        else:
            assert (got["T"] == ref["T"]).all() and (got["lifetime"] == ref["lifetime"]).all() and (got["r"] == ref["r"]).all()
    assert o.n > 1010 and w.drain_events()["removed"] > 50        # exhaust was created and swept
    w.close()


# ------------------------------------------------------------------------------------------------ chemtrail_burst, spawn
def chemtrail_cfg(**over):
    """chemtrail_burst's appcfg defaults (OON.cpp:986-995): unlimited lifetime, mass created, masses of 0.02..0.1 globe radii."""
    c = capi.EmitterCfg()
    c.v_factor, c.offset_velo_factor = 0.1, 0.2
    c.particle_lifetime, c.create_mass = -1.0, 1
    c.particle_density = 2.0
    c.velocity_divergence = 1.0
    c.particle_mass_min = 2.0 * 4.18879 * (0.02 * 5e7) ** 3
    c.particle_mass_max = 2.0 * 4.18879 * (0.1 * 5e7) ** 3
    for k, v in over.items():
        setattr(c, k, v)
    return c


def test_oracle_spawn_and_chemtrail_follow_the_reference():
    bodies, cfg = synth.make("c4_cloud_100k", n=30)
    bodies = bodies.copy()
    bodies["T"][7] = 1234.5
    bodies["v"][7] = (1e8, -2e8)
    o = make_oracle(bodies, cfg["props"])
    o.spawn(7, 4, seed=11)
    out = o.bodies()
    assert o.n == 34
    new = out[30:]
    assert (new["T"] == np.float32(1234.5)).all() and (new["v"] == bodies["v"][7]).all() and (new["lifetime"] == -1).all()
    assert (new["density"] == 1000).all()
    m0 = bodies["mass"][0]
    assert ((new["mass"] >= m0 / np.float32(9)) & (new["mass"] <= m0 * np.float32(3))).all()
    assert (np.abs(new["p"] - bodies["p"][0]) <= 0.5 * 1.5e9 * 1.0001).all()      # within 30 globe radii of the PLAYER
    r = [np.float32(pyoracle.lib().oracle_rand(11, 7 * 2 + j)) for j in range(7)]  # third body: draws 14..20
    assert new[2]["p"][0] == (r[0] * np.float32(1.5e9)) / np.float32(32767) - np.float32(1.5e9) / np.float32(2) + bodies["p"][0][0]
    assert new[2]["mass"] == m0 / np.float32(9) + (m0 * np.float32(3) - m0 / np.float32(9)) * r[6] / np.float32(32767)
    assert new[2]["color"] == (0xffffff & (int(r[4]) * int(r[5])))
    # chemtrail: positions around emitter.p + emitter.v * offset factor, within 2.5 radii
    c = chemtrail_cfg()
    assert o.chemtrail_burst(7, 6, c, seed=12) == 6
    trail = o.bodies()[34:]
    centre = bodies["p"][7] + bodies["v"][7] * np.float32(0.2)
    assert (np.abs(trail["p"] - centre) <= 2.5 * out[7]["r"] * 1.001).all()
    assert (np.abs(trail["v"] - bodies["v"][7] * np.float32(0.1)) <= 0.5 * 5e7 * 1.001).all()


@pytest.mark.gpu
@pytest.mark.parametrize("create_mass", [1, 0])
def test_device_spawn_and_chemtrail_are_bit_identical_to_the_oracle(create_mass):
    bodies, cfg = synth.make("c4_cloud_100k", n=900)
    bodies = bodies.copy()
    bodies["T"][3] = 777.0
    bodies["v"][3] = (-2e8, 5e7)
    m3 = float(bodies[3]["mass"])
    o = make_oracle(bodies, cfg["props"])
    w = gpu_world(bodies, cfg["props"], capi.MATH_EXACT)
    c = chemtrail_cfg(create_mass=create_mass, particle_lifetime=0.5, particle_mass_min=m3 * 2e-3, particle_mass_max=m3 * 6e-3)
    for burst in range(3):
        o.spawn(3, 40, seed=50 + burst); w.spawn(3, 40, seed=50 + burst)
        assert o.chemtrail_burst(3, 300, c, seed=70 + burst) == w.chemtrail_burst(3, 300, c, seed=70 + burst)
        o.step(0.033, 1); w.step(0.033, 1)
    assert w.entity_count() == o.n
    got, ref = w.download(), o.bodies()
    ok, field = bodies_equal_except_color(got, ref)
    assert ok, field
    assert (got["color"][900:940] == ref["color"][900:940]).all()      # freshly spawned bodies carry the drawn colour
    w.close()
