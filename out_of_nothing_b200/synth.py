"""Synthetic worlds for the benchmark configurations (BASELINE.json `configs`).

The reference populates worlds with ``OONApp::add_random_body_near``
(/root/reference/src/app/OON.cpp:778-803) driven by an unseeded ``rand()``; here the same
distributions are drawn from a counter-based generator (splitmix64 of (seed, body, field)) so
that every rank of a sharded run, the oracle and the tests see the identical world.

Host-side input generation only — no physics.
"""
from __future__ import annotations

import numpy as np

from .capi import BODY_F32

SEED = 20240918
CFG_GLOBE_RADIUS = 50000000.0  # World.hpp:91
PLAYER_MASS = 1e27             # test/OON.cfg:17
PLAYER_DENSITY = 550.0         # test/OON.cfg:18
DEFAULT_DENSITY = 1000.0       # Entity.hpp:45 (DENSITY_ROCK / 2)
PHYS_G = 6.6743e-11


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform(seed: int, index: np.ndarray, field: int) -> np.ndarray:
    """U[0,1) float64, a pure function of (seed, body index, field)."""
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * np.uint64(0x100000001B3) + index.astype(np.uint64) * np.uint64(64) + np.uint64(field)).astype(np.uint64)
        bits = _splitmix64(_splitmix64(key))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def radius_from_mass_and_density(mass: np.ndarray, density: np.ndarray) -> np.ndarray:
    """Entity::recalc's radius (Entity.cpp:17): r = powf((m/rho)/float(4/3 pi), float(1/3)) — the form that
    reproduces every stored radius of the KAT-1 fixture bit-exactly."""
    c = np.float32(4.0 / 3.0 * np.pi)
    v = (mass.astype(np.float32) / density.astype(np.float32) / c).astype(np.float32)
    return np.power(v.astype(np.float64), np.float64(np.float32(1.0) / np.float32(3.0))).astype(np.float32)


def world_extent(n: int) -> float:
    """Side of the spawn square: 30 globe radii up to 1000 bodies (OON.cpp:786), then grown with sqrt(N) so the
    areal density — and with it the colliding-pair fraction — stays what the reference's worlds have."""
    base = CFG_GLOBE_RADIUS * 30
    return base if n <= 1000 else base * float(np.sqrt(n / 1000.0))


def _player() -> np.ndarray:
    b = np.zeros(1, BODY_F32)
    b["lifetime"] = -1.0
    b["density"] = PLAYER_DENSITY
    b["mass"] = PLAYER_MASS
    b["color"] = 0xFFFF20
    b["free_color"] = 1
    b["gravity_immunity"] = 0          # sim/player_antigravity = false
    b["thrust"] = 0.0                  # add_thrusters(): idle thrusters mark the player
    b["r"] = radius_from_mass_and_density(b["mass"], b["density"])
    return b


def random_cloud(n: int, seed: int = SEED, extent_scale: float = 1.0) -> np.ndarray:
    """Body 0 = player at rest; bodies 1..n-1 follow add_random_body_near around it."""
    assert n >= 1
    out = np.zeros(n, BODY_F32)
    out[0] = _player()[0]
    if n == 1:
        return out
    idx = np.arange(1, n)
    L = world_extent(n) * extent_scale
    v_range = CFG_GLOBE_RADIUS * 10
    m_min, m_max = PLAYER_MASS / 9, PLAYER_MASS * 3
    b = out[1:]
    b["p"][:, 0] = (uniform(seed, idx, 0) * L - L / 2).astype(np.float32)
    b["p"][:, 1] = (uniform(seed, idx, 1) * L - L / 2).astype(np.float32)
    b["v"][:, 0] = (uniform(seed, idx, 2) * v_range - v_range / 2).astype(np.float32)
    b["v"][:, 1] = (uniform(seed, idx, 3) * v_range - v_range / 2).astype(np.float32)
    b["mass"] = (m_min + (m_max - m_min) * uniform(seed, idx, 4)).astype(np.float32)
    b["color"] = (uniform(seed, idx, 5) * 0xFFFFFF).astype(np.uint32)
    b["density"] = DEFAULT_DENSITY
    b["lifetime"] = -1.0
    b["T"] = 0.0
    b["thrust"] = 2e31
    b["r"] = radius_from_mass_and_density(b["mass"], b["density"])
    return out


def with_finite_lifetimes(bodies: np.ndarray, fraction: float, lo: float, hi: float, seed: int = SEED) -> np.ndarray:
    """Give `fraction` of the non-player bodies a lifetime ~ U(lo, hi) seconds (C3: exercises expiry + sweep)."""
    out = bodies.copy()
    idx = np.arange(1, out.shape[0])
    pick = uniform(seed, idx, 6) < fraction
    life = (lo + (hi - lo) * uniform(seed, idx, 7)).astype(np.float32)
    out["lifetime"][1:][pick] = life[pick]
    return out


def void_sphere(n: int, seed: int = SEED, gravity: float = PHYS_G) -> np.ndarray:
    """C5: bodies uniform in the annulus 0.25 L <= |p| <= 0.5 L around the player (an empty disc in the middle, as in
    doc/image/screenshot/screenshot_5000_void_sphere_zoomout_1.png), on tangential orbits for the hyperbolic law
    (F ~ G M / d  =>  v^2 = G * M_enclosed, independent of the radius)."""
    out = random_cloud(n, seed)
    if n == 1:
        return out
    idx = np.arange(1, n)
    L = world_extent(n)
    r_in, r_out = 0.25 * L, 0.5 * L
    u = uniform(seed, idx, 8)
    rad = np.sqrt(r_in * r_in + u * (r_out * r_out - r_in * r_in))
    phi = 2 * np.pi * uniform(seed, idx, 9)
    b = out[1:]
    b["p"][:, 0] = (rad * np.cos(phi)).astype(np.float32)
    b["p"][:, 1] = (rad * np.sin(phi)).astype(np.float32)
    order = np.argsort(rad)
    m_enc = np.empty(n - 1)
    m_enc[order] = PLAYER_MASS + np.cumsum(b["mass"].astype(np.float64)[order])
    speed = np.sqrt(gravity * m_enc)
    b["v"][:, 0] = (-speed * np.sin(phi)).astype(np.float32)
    b["v"][:, 1] = (speed * np.cos(phi)).astype(np.float32)
    return out


#: the benchmark configurations of BASELINE.json (C2..C5); C1 is the KAT-1 fixture itself
CONFIGS = {
    "c2_swarm_10k": dict(n=10703, friction=0.01, dt=0.033, kind="cloud"),
    "c3_dense_13k": dict(n=13503, friction=0.2, dt=0.033, kind="dense"),
    "c4_cloud_100k": dict(n=100000, friction=0.0, dt=0.033, kind="cloud"),
    "c5_void_1m": dict(n=1000000, friction=0.01, dt=0.033, kind="void"),
}


def make(config: str, n: int | None = None, seed: int = SEED) -> tuple[np.ndarray, dict]:
    """-> (bodies, props) for one of CONFIGS (optionally with a different body count)."""
    c = dict(CONFIGS[config])
    n = n or c["n"]
    if c["kind"] == "cloud":
        bodies = random_cloud(n, seed)
    elif c["kind"] == "dense":
        bodies = with_finite_lifetimes(random_cloud(n, seed, extent_scale=0.5), 0.15, 1.0, 10.0, seed)
    else:
        bodies = void_sphere(n, seed)
    props = dict(friction=c["friction"], gravity_mode=1, gravity=PHYS_G, interact_n2n=1, collision_mode=1,
                 repulsion_stiffness=0.0, loop_mode=0)
    return bodies, dict(props=props, dt=c["dt"], n=n, name=config)
