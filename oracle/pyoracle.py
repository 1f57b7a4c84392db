"""ctypes binding of the CPU oracle (oracle/liboon_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing in out_of_nothing_b200/
imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboon_oracle.so")

BODY_F32 = np.dtype(
    [
        ("gravity_immunity", "u1"), ("free_color", "u1"), ("_pad", "u1", (2,)),
        ("lifetime", "<f4"), ("r", "<f4"), ("density", "<f4"),
        ("p", "<f4", (2,)), ("v", "<f4", (2,)), ("T", "<f4"), ("color", "<u4"),
        ("mass", "<f4"), ("thrust", "<f4", (4,)),
    ]
)
assert BODY_F32.itemsize == 60


def build(force: bool = False) -> str:
    """(Re)build liboon_oracle.so when missing or older than its sources."""
    srcs = [os.path.join(_HERE, f) for f in ("oon_oracle.hpp", "oon_oracle_capi.cpp", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, u32, u64, f32, f64, i32 = (ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float,
                                       ctypes.c_double, ctypes.c_int)
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [i32]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_set_props.argtypes = [vp, f32, u32, f64, i32, f64, u32, u32]
        L.oracle_set_bodies_f32.argtypes = [vp, vp, u64, i32]
        L.oracle_body_count.restype = u64
        L.oracle_body_count.argtypes = [vp]
        L.oracle_get_bodies_f32.argtypes = [vp, vp]
        L.oracle_get_pv_f64.argtypes = [vp, vp]
        L.oracle_trace.argtypes = [vp, i32]
        L.oracle_trace_clear.argtypes = [vp]
        L.oracle_trace_counts.argtypes = [vp, vp]
        L.oracle_trace_fetch.argtypes = [vp, i32, vp]
        L.oracle_phase.argtypes = [vp, i32, f32]
        L.oracle_sweep.argtypes = [vp, u64]
        L.oracle_step.argtypes = [vp, f32, u32, i32]
        L.oracle_time_steps.restype = f64
        L.oracle_time_steps.argtypes = [vp, f32, u32, i32]
        L.oracle_radius_from_mass_and_density_f32.restype = f32
        L.oracle_radius_from_mass_and_density_f32.argtypes = [f32, f32]
        L.oracle_emit_particles.argtypes = [vp, u64, u32, vp, vp, u64]
        L.oracle_emit_particles.restype = u32
        L.oracle_chemtrail_burst.argtypes = [vp, u64, u32, vp, u64]
        L.oracle_chemtrail_burst.restype = u32
        L.oracle_spawn.argtypes = [vp, u64, u64, u32, u64]
        L.oracle_rand.argtypes = [u64, u64]
        L.oracle_rand.restype = i32
        _lib = L
    return _lib


class OracleWorld:
    """Thin object wrapper; method names follow the reference's World (World.hpp:115-128)."""

    INTRINSIC, PAIRWISE, GLOBAL, SWEEP = 0, 1, 2, 3

    def __init__(self, double: bool = False):
        self._h = lib().oracle_create(1 if double else 0)
        self.double = double

    def __del__(self):
        try:
            if self._h:
                lib().oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_props(self, friction=0.03, gravity_mode=1, gravity=6.6743e-11, interact_n2n=False,
                  repulsion_stiffness=0.0, collision_mode=1, loop_mode=0):
        lib().oracle_set_props(self._h, friction, gravity_mode, gravity, int(bool(interact_n2n)),
                               repulsion_stiffness, collision_mode, loop_mode)

    def set_bodies(self, aos: np.ndarray, recalc: bool = False):
        aos = np.ascontiguousarray(aos)
        assert aos.dtype.itemsize == 60
        lib().oracle_set_bodies_f32(self._h, aos.ctypes.data, aos.shape[0], int(recalc))

    @property
    def n(self) -> int:
        return int(lib().oracle_body_count(self._h))

    def bodies(self) -> np.ndarray:
        out = np.zeros(self.n, BODY_F32)
        if self.n:
            lib().oracle_get_bodies_f32(self._h, out.ctypes.data)
        return out

    def pv64(self) -> np.ndarray:
        out = np.zeros((self.n, 4), np.float64)
        if self.n:
            lib().oracle_get_pv_f64(self._h, out.ctypes.data)
        return out

    def trace(self, on: bool = True):
        lib().oracle_trace(self._h, int(on))

    def trace_clear(self):
        lib().oracle_trace_clear(self._h)

    def trace_fetch(self):
        """-> dict(collisions=[k,2] (target,source), touches=[k,2], expired=[k], terminated=[k], pairs_visited)."""
        c = (ctypes.c_uint64 * 5)()
        lib().oracle_trace_counts(self._h, c)
        out = {}
        for which, (name, width) in enumerate((("collisions", 2), ("touches", 2), ("expired", 1), ("terminated", 1))):
            k = int(c[which])
            a = np.zeros((k, width), np.uint32)
            if k:
                lib().oracle_trace_fetch(self._h, which, a.ctypes.data)
            out[name] = a if width == 2 else a[:, 0]
        out["pairs_visited"] = int(c[4])
        return out

    def phase(self, which: int, dt: float):
        lib().oracle_phase(self._h, which, dt)

    def update_intrinsic(self, dt): self.phase(self.INTRINSIC, dt)
    def update_pairwise(self, dt): self.phase(self.PAIRWISE, dt)
    def update_global(self, dt): self.phase(self.GLOBAL, dt)
    def sweep_expired(self, player_ndx: int = 0): lib().oracle_sweep(self._h, player_ndx)

    #: field order of the 15-float configuration oracle_emit_particles takes (== Emitter::Config, Emitter.hpp:24-37)
    EMITTER_FIELDS = ("eject_velocity_x", "eject_velocity_y", "eject_offset_x", "eject_offset_y", "v_factor", "offset_velo_factor",
                      "particle_lifetime", "create_mass", "particle_density", "position_divergence_x", "position_divergence_y",
                      "velocity_divergence", "particle_mass_min", "particle_mass_max", "color")

    def _emitter_floats(self, cfg) -> np.ndarray:
        c = np.zeros(15, np.float32)
        for i, name in enumerate(self.EMITTER_FIELDS[:14]):
            c[i] = float(getattr(cfg, name))
        c[14:15].view(np.uint32)[0] = int(getattr(cfg, "color"))
        return c

    def emit_particles(self, emitter_id: int, n: int, cfg, nozzles=None, seed: int = 0) -> int:
        """Emitter::emit_particles (Emitter.cpp:22-92); cfg = any object with the EMITTER_FIELDS attributes."""
        c = self._emitter_floats(cfg)
        nz = None if nozzles is None else np.ascontiguousarray(nozzles, np.float32)
        return int(lib().oracle_emit_particles(self._h, emitter_id, n, c.ctypes.data, nz.ctypes.data if nz is not None else None, seed))

    def chemtrail_burst(self, emitter_id: int, n: int, cfg, seed: int = 0) -> int:
        """OONApp::chemtrail_burst (OON.cpp:984-1029)."""
        c = self._emitter_floats(cfg)
        return int(lib().oracle_chemtrail_burst(self._h, emitter_id, n, c.ctypes.data, seed))

    def spawn(self, parent_id: int, n: int, seed: int = 0, player_id: int = 0):
        """OONApp::spawn (OON.cpp:842-861) over add_random_body_near (OON.cpp:778-803)."""
        lib().oracle_spawn(self._h, parent_id, player_id, n, seed)

    def step(self, dt: float, nsteps: int = 1, sweep: bool = True):
        lib().oracle_step(self._h, dt, nsteps, int(sweep))

    def time_steps(self, dt: float, nsteps: int, sweep: bool = True) -> float:
        return float(lib().oracle_time_steps(self._h, dt, nsteps, int(sweep)))


def radius_from_mass_and_density(mass, density) -> float:
    return float(lib().oracle_radius_from_mass_and_density_f32(mass, density))
