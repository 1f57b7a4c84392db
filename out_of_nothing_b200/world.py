"""Python host mirror of the reference's ``OON::Model::World`` over the C ABI.

Same names and argument meaning as /root/reference/src/app/model/World.hpp:82-158
(``props``, ``add_body``, ``remove_body``, ``update_intrinsic/_pairwise/_global``,
``loop_mode``/``set_loop_mode``) plus the caller-side sweep of OON.cpp:1089-1095, so that
parity tests read like the reference's own regression runs.  All physics happens in
liboon_gpu.so on the device; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import capi
from .capi import BODY_F32, EmitterCfg, Events, OonError, Props


def recalc_radii(bodies: np.ndarray) -> np.ndarray:
    """``Entity::recalc()`` (Entity.cpp:14-20) for a whole table: r = radius_from_mass_and_density(mass, density).
    (The colour half of recalc() cannot be restated: Phys::T_to_RGB_and_BV is absent from the reference tree.)"""
    from .synth import radius_from_mass_and_density
    out = np.ascontiguousarray(bodies).copy()
    out["r"] = radius_from_mass_and_density(out["mass"], out["density"])
    return out


class Properties:
    """Live view of ``World::props`` (World.hpp:59-69): attribute writes go straight to the device handle."""

    _names = ("friction", "gravity_mode", "gravity", "interact_n2n", "collision_mode", "repulsion_stiffness", "loop_mode")

    def __init__(self, world: "World"):
        object.__setattr__(self, "_w", world)

    def __getattr__(self, name):
        if name not in self._names:
            raise AttributeError(name)
        p = self._w._get_props()
        v = getattr(p, name)
        return bool(v) if name == "interact_n2n" else v

    def __setattr__(self, name, value):
        if name not in self._names:
            raise AttributeError(name)
        p = self._w._get_props()
        setattr(p, name, int(value) if name in ("gravity_mode", "interact_n2n", "collision_mode", "loop_mode") else value)
        self._w._set_props(p)


class World:
    """One world on one GPU, one world over several GPUs of this process (``devices=[...]``), or one shard of a world
    sharded over ``nranks`` processes."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, nccl_id: Optional[bytes] = None, devices=None):
        self._L = capi.load()
        self._h = ctypes.c_void_p()
        if devices is not None:                              # single-process multi-GPU handle (oon_create_multi)
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self._L.oon_create_multi(arr, len(devices), ctypes.byref(self._h))
            device, rank, nranks = devices[0], 0, 1           # to its caller the group is one rank that owns the whole table
        elif nranks == 1:
            rc = self._L.oon_create(device, ctypes.byref(self._h))
        else:
            rc = self._L.oon_create_sharded(device, rank, nranks, nccl_id, ctypes.byref(self._h))
        if rc:
            raise OonError(rc, (self._L.oon_last_error(None) or b"").decode())
        self.props = Properties(self)
        self.rank, self.nranks, self.device = rank, nranks, device

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc:
            raise OonError(rc, (self._L.oon_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.oon_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _get_props(self) -> Props:
        p = Props()
        self._check(self._L.oon_get_props(self._h, ctypes.byref(p)))
        return p

    def _set_props(self, p: Props):
        self._check(self._L.oon_set_props(self._h, ctypes.byref(p)))

    # ------------------------------------------------------------------ knobs
    def set_props(self, **kw):
        p = self._get_props()
        for k, v in kw.items():
            if k not in Properties._names:
                raise AttributeError(k)
            setattr(p, k, int(v) if k in ("gravity_mode", "interact_n2n", "collision_mode", "loop_mode") else v)
        self._set_props(p)

    def loop_mode(self) -> int:
        return self._get_props().loop_mode

    def set_loop_mode(self, mode: int):
        self.set_props(loop_mode=mode)

    @property
    def math_mode(self) -> int:
        m = ctypes.c_int()
        self._check(self._L.oon_get_math_mode(self._h, ctypes.byref(m)))
        return m.value

    @math_mode.setter
    def math_mode(self, mode: int):
        self._check(self._L.oon_set_math_mode(self._h, mode))

    # ------------------------------------------------------------------ body table
    def upload(self, bodies: np.ndarray):
        bodies = np.ascontiguousarray(bodies)
        assert bodies.dtype.itemsize == 60, "expected 60-byte float-build records"
        self._check(self._L.oon_upload(self._h, bodies.ctypes.data, bodies.shape[0]))

    def upload_ptr(self, ptr: int, n: int):
        """Upload from a raw host pointer (e.g. pinned memory owned by the caller)."""
        self._check(self._L.oon_upload(self._h, ptr, n))

    def download(self) -> np.ndarray:
        n = self.entity_count()
        out = np.zeros(n, BODY_F32)
        got = ctypes.c_uint64()
        self._check(self._L.oon_download(self._h, out.ctypes.data, n, ctypes.byref(got)))
        return out[: got.value]

    def download_ptr(self, ptr: int, capacity: int) -> int:
        got = ctypes.c_uint64()
        self._check(self._L.oon_download(self._h, ptr, capacity, ctypes.byref(got)))
        return got.value

    # shard-local I/O (sharded worlds): only this rank's block crosses the bus, no collective on the data path
    def local_range(self, n_total: int):
        """(first, count) of the block this rank owns in a freshly uploaded world of n_total bodies."""
        first, count = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self._L.oon_shard_range(n_total, self.rank, self.nranks, ctypes.byref(first), ctypes.byref(count)))
        return first.value, count.value

    def upload_local(self, block: np.ndarray, n_total: int):
        block = np.ascontiguousarray(block)
        assert block.dtype.itemsize == 60
        self._check(self._L.oon_upload_local(self._h, block.ctypes.data if block.shape[0] else None, block.shape[0], n_total))

    def upload_local_ptr(self, ptr: int, count: int, n_total: int):
        self._check(self._L.oon_upload_local(self._h, ptr, count, n_total))

    def download_local(self):
        """-> (first global index, this rank's bodies).  Collective on a multi-process world whose bodies can expire."""
        first, got = ctypes.c_uint64(), ctypes.c_uint64()
        rc = self._L.oon_download_local(self._h, None, 0, ctypes.byref(first), ctypes.byref(got))     # how many? (ragged blocks after sweeps)
        if rc not in (capi.OK, capi.ERR_CAPACITY):
            self._check(rc)
        out = np.zeros(got.value, BODY_F32)
        if got.value:
            self._check(self._L.oon_download_local(self._h, out.ctypes.data, got.value, ctypes.byref(first), ctypes.byref(got)))
        return first.value, out[: got.value]

    def download_local_ptr(self, ptr: int, capacity: int):
        first, got = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self._L.oon_download_local(self._h, ptr, capacity, ctypes.byref(first), ctypes.byref(got)))
        return first.value, got.value

    def set_player_index(self, ndx: int):
        self._check(self._L.oon_set_player_index(self._h, ndx))

    def download_render(self) -> np.ndarray:
        n = self.entity_count()
        out = np.zeros((n, 3), np.float32)
        got = ctypes.c_uint64()
        self._check(self._L.oon_download_render(self._h, out.ctypes.data, n, ctypes.byref(got)))
        return out[: got.value]

    def download_render_rgb(self):
        """-> ({x, y, r} per body, colour word per body): everything render_scene reads (OONMainDisplay_sfml.cpp:181-211)."""
        n = self.entity_count()
        xyr, col = np.zeros((n, 3), np.float32), np.zeros(n, np.uint32)
        got = ctypes.c_uint64()
        self._check(self._L.oon_download_render_rgb(self._h, xyr.ctypes.data, col.ctypes.data, n, ctypes.byref(got)))
        return xyr[: got.value], col[: got.value]

    def device_count(self) -> int:
        n = ctypes.c_int()
        self._check(self._L.oon_device_count(self._h, ctypes.byref(n)))
        return n.value

    def download_render_begin(self, ptr: int, capacity: int):
        """Start the render read-back ({x, y, r} per body) into pinned host memory at ``ptr``; returns at once."""
        self._check(self._L.oon_download_render_begin(self._h, ptr, capacity))

    def download_render_end(self) -> int:
        """Wait for the read-back started by download_render_begin; returns the number of bodies it holds."""
        n = ctypes.c_uint64()
        self._check(self._L.oon_download_render_end(self._h, ctypes.byref(n)))
        return n.value

    def entity_count(self) -> int:
        n = ctypes.c_uint64()
        self._check(self._L.oon_body_count(self._h, ctypes.byref(n)))
        return n.value

    def add_body(self, body, recalc: bool = True) -> int:
        """World::add_body (World.cpp:55-70): append + ``recalc()`` — the radius is recomputed from mass and density
        (Entity.cpp:17) like the reference does for every body it adds or loads; ``recalc=False`` keeps the stored
        radius (raw table edits).  Returns the new body's index."""
        rec = np.zeros(1, BODY_F32)
        rec[0] = body
        n = self.entity_count()
        self.add_bodies(rec, recalc=recalc)
        return n

    def add_bodies(self, bodies: np.ndarray, recalc: bool = True):
        bodies = recalc_radii(bodies) if recalc else np.ascontiguousarray(bodies)
        self._check(self._L.oon_add_bodies(self._h, bodies.ctypes.data, bodies.shape[0]))

    def load(self, snap, recalc: bool = True):
        """World::load (World_SaveLoad.cpp:125-165): take the knobs of a snapshot header and add its bodies one by one
        through add_body(), i.e. WITH ``recalc()`` — v0.0.1 snapshots store radii of an older convention and the
        reference replaces them on load.  ``snap`` = a ``snapshot.Snapshot`` or a path."""
        from . import snapshot as _snapshot
        if not hasattr(snap, "bodies"):
            snap = _snapshot.load(snap)
        self.set_props(friction=snap.friction, gravity_mode=snap.gravity_mode, gravity=snap.gravity, interact_n2n=snap.interact_n2n)
        bodies = snap.as_f32()                     # a double-build file is narrowed like oon_upload_f64 does
        self.upload(recalc_radii(bodies) if recalc else bodies)
        return snap

    def remove_body(self, ndx: int):
        """World::remove_body (World.cpp:72-78): erase, later indices shift down."""
        self._check(self._L.oon_remove_body(self._h, ndx))

    def entity(self, ndx: int) -> np.ndarray:
        rec = np.zeros(1, BODY_F32)
        self._check(self._L.oon_get_body(self._h, ndx, rec.ctypes.data))
        return rec[0]

    def set_entity(self, ndx: int, body):
        rec = np.zeros(1, BODY_F32)
        rec[0] = body
        self._check(self._L.oon_set_body(self._h, ndx, rec.ctypes.data))

    # ------------------------------------------------------------------ the hot path
    def update_intrinsic(self, dt: float): self._check(self._L.oon_update_intrinsic(self._h, dt))
    def update_pairwise(self, dt: float): self._check(self._L.oon_update_pairwise(self._h, dt))
    def update_global(self, dt: float): self._check(self._L.oon_update_global(self._h, dt))
    def sweep_expired(self, player_ndx: int = 0): self._check(self._L.oon_sweep_expired(self._h, player_ndx))

    def step(self, dt: float, nsteps: int = 1):
        """nsteps frames of OONApp::updates_for_next_frame's model part (OON.cpp:1085-1095)."""
        self._check(self._L.oon_step(self._h, dt, nsteps))

    def step_profiled(self, dt: float, nsteps: int = 1):
        ms = (ctypes.c_float * 4)()
        launches = (ctypes.c_uint32 * 4)()
        self._check(self._L.oon_step_profiled(self._h, dt, nsteps, ms, launches))
        names = ("interact", "integrate", "exchange", "sweep")
        return {k: float(ms[i]) for i, k in enumerate(names)}, {k: int(launches[i]) for i, k in enumerate(names)}

    def bench_frames(self, dt: float, nsteps: int, flush_bytes: int = 0) -> float:
        """nsteps frames, L2 flushed by a device memset before each, only the frames timed (CUDA events); -> total ms."""
        ms = ctypes.c_double()
        self._check(self._L.oon_bench_frames(self._h, dt, nsteps, flush_bytes, ctypes.byref(ms)))
        return ms.value

    def bench_frames_profiled(self, dt: float, nsteps: int, flush_bytes: int = 0):
        """bench_frames with per-kernel-class event pairs as well -> (total ms, {class: ms}, {class: launches})."""
        total = ctypes.c_double()
        ms = (ctypes.c_float * 4)()
        launches = (ctypes.c_uint32 * 4)()
        self._check(self._L.oon_bench_frames_profiled(self._h, dt, nsteps, flush_bytes, ctypes.byref(total), ms, launches))
        names = ("interact", "integrate", "exchange", "sweep")
        return total.value, {k: float(ms[i]) for i, k in enumerate(names)}, {k: int(launches[i]) for i, k in enumerate(names)}

    def synchronize(self): self._check(self._L.oon_synchronize(self._h))

    def drain_events(self) -> dict:
        ev = Events()
        self._check(self._L.oon_drain_events(self._h, ctypes.byref(ev)))
        return {n: int(getattr(ev, n)) for n, _ in Events._fields_}

    # ------------------------------------------------------------------ measurement / trace
    def timer_start(self): self._check(self._L.oon_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = ctypes.c_float()
        self._check(self._L.oon_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def emitter_config(self, **kw) -> EmitterCfg:
        """``Emitter::Config`` with the reference's defaults (Emitter.hpp:24-37), overridden by keyword."""
        c = EmitterCfg()
        self._check(self._L.oon_emitter_cfg_default(ctypes.byref(c)))
        for k, v in kw.items():
            if k in ("eject_velocity", "eject_offset", "position_divergence"):
                setattr(c, k + "_x", v[0]); setattr(c, k + "_y", v[1])
            elif k in ("create_mass", "color"):
                setattr(c, k, int(v))
            else:
                if not hasattr(c, k):
                    raise AttributeError(k)
                setattr(c, k, v)
        return c

    def emit_particles(self, emitter_id: int, n: int, cfg: EmitterCfg, nozzles=None, seed: int = 0) -> int:
        """``Emitter::emit_particles`` (Emitter.cpp:22-92) on the device; returns the number of particles appended."""
        nz = None
        if nozzles is not None:
            nz = np.ascontiguousarray(nozzles, np.float32)
            assert nz.shape == (n, 2)
        out = ctypes.c_uint32()
        self._check(self._L.oon_emit_particles(self._h, emitter_id, n, ctypes.byref(cfg), nz.ctypes.data if nz is not None else None,
                                               seed, ctypes.byref(out)))
        return out.value

    def chemtrail_burst(self, emitter_id: int, n: int, cfg: EmitterCfg, seed: int = 0) -> int:
        """``OONApp::chemtrail_burst`` (OON.cpp:984-1029) on the device; returns the number of particles appended."""
        out = ctypes.c_uint32()
        self._check(self._L.oon_chemtrail_burst(self._h, emitter_id, n, ctypes.byref(cfg), seed, ctypes.byref(out)))
        return out.value

    def spawn(self, parent_id: int, n: int, seed: int = 0, player_id: int = 0):
        """``OONApp::spawn`` (OON.cpp:842-861): n random bodies near the player, inheriting T and v from the parent."""
        self._check(self._L.oon_spawn(self._h, parent_id, player_id, n, seed))

    def exchange_kind(self) -> str:
        """How a sharded world moves its exchange tiles: 'none' (one GPU), 'nccl_allgather' or 'peer_stores' (NVLink peer memory)."""
        k = ctypes.c_int()
        self._check(self._L.oon_exchange_kind(self._h, ctypes.byref(k)))
        return ("none", "nccl_allgather", "peer_stores")[k.value]

    def debug_time_interact(self, dt: float, fraction: int, part: int = 0) -> float:
        """ms of one FAST all-pairs launch over one of `fraction` parts of the targets; the world is left unchanged."""
        ms = ctypes.c_float()
        self._check(self._L.oon_debug_time_interact(self._h, dt, fraction, part, ctypes.byref(ms)))
        return ms.value

    def debug_trace_interact(self, dt: float, fraction: int, part: int = 0, capacity: int = 1 << 16):
        """-> (ms, trace[n_ctas, 3] = {SM id | tiles << 32, start ns, end ns}) of the launch debug_time_interact times."""
        out = np.zeros((capacity, 3), np.uint64)
        n, ms = ctypes.c_uint64(), ctypes.c_float()
        self._check(self._L.oon_debug_trace_interact(self._h, dt, fraction, part, out.ctypes.data, capacity, ctypes.byref(n), ctypes.byref(ms)))
        return ms.value, out[: n.value]

    def launch_count(self) -> int:
        n = ctypes.c_uint64()
        self._check(self._L.oon_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def debug_capture_pairs(self, capacity: int): self._check(self._L.oon_debug_capture_pairs(self._h, capacity))

    def debug_fetch_pairs(self, which: int, capacity: int = 1 << 22) -> np.ndarray:
        out = np.zeros((capacity, 2), np.uint32)
        n = ctypes.c_uint64()
        self._check(self._L.oon_debug_fetch_pairs(self._h, which, out.ctypes.data, capacity, ctypes.byref(n)))
        return out[: n.value].copy()

    def debug_fetch_kick(self) -> np.ndarray:
        n = self.entity_count()
        out = np.zeros((n, 2), np.float32)
        got = ctypes.c_uint64()
        self._check(self._L.oon_debug_fetch_kick(self._h, out.ctypes.data, n, ctypes.byref(got)))
        return out[: got.value]

    def debug_fetch_removed(self, capacity: int = 1 << 20) -> np.ndarray:
        out = np.zeros(capacity, np.uint32)
        n = ctypes.c_uint64()
        self._check(self._L.oon_debug_fetch_removed(self._h, out.ctypes.data, capacity, ctypes.byref(n)))
        return out[: n.value].copy()
