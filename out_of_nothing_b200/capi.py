"""ctypes binding of the C-ABI library (include/oon_gpu.h -> out_of_nothing_b200/lib/liboon_gpu.so).

This is the binding a Python host would add; tests and bench.py call the product through it.
There is deliberately no fallback: if the shared library is missing or no sm_100 device is
visible, loading / ``World()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OON_GPU_LIB") or os.path.join(_HERE, "lib", "liboon_gpu.so")   # override: tuning variants only

OK, ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_NO_DEVICE, ERR_CAPACITY, ERR_UNSUPPORTED = range(7)
GRAVITY_OFF, GRAVITY_HYPERBOLIC, GRAVITY_REALISTIC, GRAVITY_EXPERIMENTAL = range(4)
COLLISION_OFF, COLLISION_GLIDE_THROUGH = 0, 1
LOOP_HALF_MATRIX, LOOP_FULL_MATRIX = 0, 1
MATH_FAST, MATH_EXACT = 0, 1
NCCL_ID_BYTES = 128

#: numpy view of ``oon_body_f32`` (== the reference's 60-byte float-build Entity)
BODY_F32 = np.dtype(
    [
        ("gravity_immunity", "u1"), ("free_color", "u1"), ("_pad", "u1", (2,)),
        ("lifetime", "<f4"), ("r", "<f4"), ("density", "<f4"),
        ("p", "<f4", (2,)), ("v", "<f4", (2,)), ("T", "<f4"), ("color", "<u4"),
        ("mass", "<f4"), ("thrust", "<f4", (4,)),
    ]
)
assert BODY_F32.itemsize == 60


class Props(ctypes.Structure):
    """``oon_props`` == World::Properties + loop mode (World.hpp:59-69)."""

    _fields_ = [
        ("friction", ctypes.c_float),
        ("gravity_mode", ctypes.c_uint32),
        ("gravity", ctypes.c_double),
        ("interact_n2n", ctypes.c_uint32),
        ("collision_mode", ctypes.c_uint32),
        ("repulsion_stiffness", ctypes.c_double),
        ("loop_mode", ctypes.c_uint32),
        ("_reserved", ctypes.c_uint32),
    ]


class EmitterCfg(ctypes.Structure):
    """``oon_emitter_cfg`` == Emitter::Config (Emitter.hpp:24-37), float build."""

    _fields_ = [
        ("eject_velocity_x", ctypes.c_float), ("eject_velocity_y", ctypes.c_float),
        ("eject_offset_x", ctypes.c_float), ("eject_offset_y", ctypes.c_float),
        ("v_factor", ctypes.c_float), ("offset_velo_factor", ctypes.c_float),
        ("particle_lifetime", ctypes.c_float), ("create_mass", ctypes.c_uint32),
        ("particle_density", ctypes.c_float),
        ("position_divergence_x", ctypes.c_float), ("position_divergence_y", ctypes.c_float),
        ("velocity_divergence", ctypes.c_float),
        ("particle_mass_min", ctypes.c_float), ("particle_mass_max", ctypes.c_float),
        ("color", ctypes.c_uint32),
    ]


assert ctypes.sizeof(EmitterCfg) == 60


class Events(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("steps", "collisions", "touches", "player_touches", "terminated", "removed",
                                                "warp_tiles", "warp_tiles_checked")]


#: every symbol include/oon_gpu.h declares (tests check the library exports exactly these)
EXPORTS = [
    "oon_abi_version", "oon_create", "oon_nccl_unique_id", "oon_create_sharded", "oon_create_multi", "oon_device_count", "oon_destroy", "oon_last_error",
    "oon_props_default", "oon_set_props", "oon_get_props", "oon_set_math_mode", "oon_get_math_mode",
    "oon_upload", "oon_upload_f64", "oon_shard_range", "oon_upload_local", "oon_download_local", "oon_set_player_index", "oon_download", "oon_download_f64", "oon_download_render", "oon_download_render_rgb", "oon_download_render_begin", "oon_download_render_end", "oon_body_count",
    "oon_add_bodies", "oon_remove_body", "oon_get_body", "oon_set_body", "oon_emitter_cfg_default", "oon_emit_particles", "oon_chemtrail_burst", "oon_spawn",
    "oon_step", "oon_update_intrinsic", "oon_update_pairwise", "oon_update_global", "oon_sweep_expired",
    "oon_synchronize", "oon_drain_events",
    "oon_timer_start", "oon_timer_stop", "oon_step_profiled", "oon_bench_frames", "oon_bench_frames_profiled", "oon_launch_count", "oon_exchange_kind",
    "oon_measure_fp32_peak", "oon_measure_copy_bandwidth",
    "oon_debug_capture_pairs", "oon_debug_fetch_pairs", "oon_debug_fetch_kick", "oon_debug_fetch_removed", "oon_debug_time_interact", "oon_debug_trace_interact", "oon_debug_ieee_selftest",
]

_lib = None


class OonError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("oon error %d: %s" % (code, message))
        self.code = code


def load():
    """Load liboon_gpu.so (raises if it has not been built: run ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OonError(ERR_NO_DEVICE, "CUDA library %s is missing; build it first (no CPU fallback exists)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64, f32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float, ctypes.c_int
    P = ctypes.POINTER
    L.oon_abi_version.restype = i32
    L.oon_create.argtypes = [i32, P(vp)]
    L.oon_nccl_unique_id.argtypes = [vp]
    L.oon_create_sharded.argtypes = [i32, i32, i32, vp, P(vp)]
    L.oon_create_multi.argtypes = [P(i32), i32, P(vp)]
    L.oon_device_count.argtypes = [vp, P(i32)]
    L.oon_destroy.argtypes = [vp]
    L.oon_last_error.restype = ctypes.c_char_p
    L.oon_last_error.argtypes = [vp]
    L.oon_props_default.argtypes = [P(Props)]
    L.oon_set_props.argtypes = [vp, P(Props)]
    L.oon_get_props.argtypes = [vp, P(Props)]
    L.oon_set_math_mode.argtypes = [vp, i32]
    L.oon_get_math_mode.argtypes = [vp, P(i32)]
    L.oon_upload.argtypes = [vp, vp, u64]
    L.oon_upload_f64.argtypes = [vp, vp, u64]
    L.oon_download.argtypes = [vp, vp, u64, P(u64)]
    L.oon_shard_range.argtypes = [u64, i32, i32, P(u64), P(u64)]
    L.oon_upload_local.argtypes = [vp, vp, u64, u64]
    L.oon_download_local.argtypes = [vp, vp, u64, P(u64), P(u64)]
    L.oon_set_player_index.argtypes = [vp, u64]
    L.oon_download_f64.argtypes = [vp, vp, u64, P(u64)]
    L.oon_download_render.argtypes = [vp, vp, u64, P(u64)]
    L.oon_download_render_rgb.argtypes = [vp, vp, vp, u64, P(u64)]
    L.oon_body_count.argtypes = [vp, P(u64)]
    L.oon_add_bodies.argtypes = [vp, vp, u64]
    L.oon_remove_body.argtypes = [vp, u64]
    L.oon_get_body.argtypes = [vp, u64, vp]
    L.oon_set_body.argtypes = [vp, u64, vp]
    L.oon_emitter_cfg_default.argtypes = [P(EmitterCfg)]
    L.oon_emit_particles.argtypes = [vp, u64, u32, P(EmitterCfg), vp, u64, P(u32)]
    L.oon_chemtrail_burst.argtypes = [vp, u64, u32, P(EmitterCfg), u64, P(u32)]
    L.oon_spawn.argtypes = [vp, u64, u64, u32, u64]
    L.oon_download_render_begin.argtypes = [vp, vp, u64]
    L.oon_download_render_end.argtypes = [vp, P(u64)]
    L.oon_step.argtypes = [vp, f32, u32]
    L.oon_update_intrinsic.argtypes = [vp, f32]
    L.oon_update_pairwise.argtypes = [vp, f32]
    L.oon_update_global.argtypes = [vp, f32]
    L.oon_sweep_expired.argtypes = [vp, u64]
    L.oon_synchronize.argtypes = [vp]
    L.oon_drain_events.argtypes = [vp, P(Events)]
    L.oon_timer_start.argtypes = [vp]
    L.oon_timer_stop.argtypes = [vp, P(f32)]
    L.oon_step_profiled.argtypes = [vp, f32, u32, P(f32), P(u32)]
    L.oon_bench_frames.argtypes = [vp, f32, u32, u64, P(ctypes.c_double)]
    L.oon_bench_frames_profiled.argtypes = [vp, f32, u32, u64, P(ctypes.c_double), P(f32), P(u32)]
    L.oon_launch_count.argtypes = [vp, P(u64)]
    L.oon_exchange_kind.argtypes = [vp, P(ctypes.c_int)]
    L.oon_measure_fp32_peak.argtypes = [i32, P(ctypes.c_double)]
    L.oon_measure_copy_bandwidth.argtypes = [i32, P(ctypes.c_double)]
    L.oon_debug_capture_pairs.argtypes = [vp, u64]
    L.oon_debug_fetch_pairs.argtypes = [vp, i32, vp, u64, P(u64)]
    L.oon_debug_fetch_kick.argtypes = [vp, vp, u64, P(u64)]
    L.oon_debug_fetch_removed.argtypes = [vp, vp, u64, P(u64)]
    L.oon_debug_time_interact.argtypes = [vp, f32, u32, u32, P(f32)]
    L.oon_debug_ieee_selftest.argtypes = [i32, u64, u64, P(u64)]
    L.oon_debug_trace_interact.argtypes = [vp, f32, u32, u32, vp, u64, P(u64), P(f32)]
    for name in EXPORTS:
        if name != "oon_last_error":
            getattr(L, name).restype = i32
    _lib = L
    return L


def nccl_unique_id() -> bytes:
    L = load()
    buf = ctypes.create_string_buffer(NCCL_ID_BYTES)
    rc = L.oon_nccl_unique_id(buf)
    if rc:
        raise OonError(rc, (L.oon_last_error(None) or b"").decode())
    return buf.raw


def measure_fp32_peak(device: int = 0) -> float:
    L = load()
    out = ctypes.c_double()
    rc = L.oon_measure_fp32_peak(device, ctypes.byref(out))
    if rc:
        raise OonError(rc, (L.oon_last_error(None) or b"").decode())
    return out.value


def measure_copy_bandwidth(device: int = 0) -> float:
    L = load()
    out = ctypes.c_double()
    rc = L.oon_measure_copy_bandwidth(device, ctypes.byref(out))
    if rc:
        raise OonError(rc, (L.oon_last_error(None) or b"").decode())
    return out.value


def ieee_selftest(device: int = 0, n: int = 1 << 26, seed: int = 1) -> int:
    """mismatches between the EXACT kernel's written-out sqrt.rn / div.rn fast paths and the intrinsics over n operand sets"""
    L = load()
    out = ctypes.c_uint64()
    rc = L.oon_debug_ieee_selftest(device, n, seed, ctypes.byref(out))
    if rc:
        raise OonError(rc, (L.oon_last_error(None) or b"").decode())
    return out.value


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data
