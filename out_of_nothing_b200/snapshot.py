"""Snapshot (``*.state`` / ``*.save``) reader and writer.

Host-side test/bench utility: decodes the reference's world snapshot format so
fixtures can be fed to the C-ABI library and to the oracle, and encodes worlds
back into the same format.  No physics here.

Format follows the reference's writer/loader
  - /root/reference/src/app/model/World_SaveLoad.cpp:36-72 (save), :80-165 (load)
  - /root/reference/src/app/model/Entity_SaveLoad.cpp:23-49 (save), :53-77 (load)
  - record layout = raw ``Entity`` bytes, /root/reference/src/app/model/Entity.hpp:25-98
and SURVEY.md Appendix A (layout decoded from the fixtures).

Text header ``name = value`` lines, a mandatory ``- - -`` separator, then one
line per body: ``<ndx> : "<raw struct bytes with " and \\ backslash-escaped>"``.
The whole file may be wrapped in one zstd frame (engine-side option
``save_compressed``); it is inflated with the system ``libzstd.so.1`` via ctypes.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import dataclasses
from typing import Dict, Optional

import numpy as np

ZSTD_MAGIC = b"\x28\xb5\x2f\xfd"

#: float build of ``Entity`` — 60 bytes (36 of the 37 fixtures), == ``oon_body_f32`` in include/oon_gpu.h
BODY_F32 = np.dtype(
    [
        ("gravity_immunity", "u1"),
        ("free_color", "u1"),
        ("_pad", "u1", (2,)),
        ("lifetime", "<f4"),
        ("r", "<f4"),
        ("density", "<f4"),
        ("p", "<f4", (2,)),
        ("v", "<f4", (2,)),
        ("T", "<f4"),
        ("color", "<u4"),
        ("mass", "<f4"),
        ("thrust", "<f4", (4,)),  # up, down, left, right
    ]
)
assert BODY_F32.itemsize == 60

#: double build of ``Entity`` — 96 bytes, == ``oon_body_f64``
BODY_F64 = np.dtype(
    [
        ("gravity_immunity", "u1"),
        ("free_color", "u1"),
        ("_pad", "u1", (6,)),
        ("lifetime", "<f8"),
        ("r", "<f8"),
        ("density", "<f8"),
        ("p", "<f8", (2,)),
        ("v", "<f8", (2,)),
        ("T", "<f4"),
        ("color", "<u4"),
        ("mass", "<f8"),
        ("thrust", "<f4", (4,)),
    ]
)
assert BODY_F64.itemsize == 96

#: ``Math::unset<float>()`` as stored in every non-player body's thruster fields
#: (value read from the fixtures; /root/reference/src/app/model/Entity.hpp:70-73,87)
THRUST_UNSET = np.float32(2e31)

MODEL_VERSION = "0.1.5"  # /root/reference/src/app/model/World.hpp:36

# GravityMode, /root/reference/src/app/model/World.hpp:42-50
GRAVITY_OFF, GRAVITY_HYPERBOLIC, GRAVITY_REALISTIC, GRAVITY_EXPERIMENTAL = 0, 1, 2, 3
PHYS_G = 6.6743e-11  # runtime default written by a 0.0.1 -> 0.1.x re-save (kat3_end header)


@dataclasses.dataclass
class Snapshot:
    """One decoded world: global properties + the body table (AoS, host order)."""

    version: str
    friction: float
    interact_n2n: bool
    gravity_mode: int
    gravity: float
    bodies: np.ndarray  # BODY_F32 or BODY_F64
    header: Dict[str, str] = dataclasses.field(default_factory=dict)
    compressed: bool = False

    @property
    def n(self) -> int:
        return int(self.bodies.shape[0])

    def as_f32(self) -> np.ndarray:
        """Body table in the 60-byte float layout (narrowing if the file was a double build)."""
        if self.bodies.dtype == BODY_F32:
            return self.bodies.copy()
        out = np.zeros(self.n, BODY_F32)
        for name in BODY_F32.names:
            if name != "_pad":
                out[name] = self.bodies[name]
        return out


def _version_tuple(v: str):
    return tuple(int(x) for x in v.strip().split("."))


# --------------------------------------------------------------------------- zstd
_zstd = None


def _libzstd():
    global _zstd
    if _zstd is None:
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        lib = ctypes.CDLL(name)
        lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        _zstd = lib
    return _zstd


def zstd_decompress(data: bytes) -> bytes:
    lib = _libzstd()
    size = lib.ZSTD_getFrameContentSize(data, len(data))
    if size in (2**64 - 1, 2**64 - 2):  # ERROR / UNKNOWN
        size = 64 * len(data) + (1 << 20)
    buf = ctypes.create_string_buffer(int(size))
    got = lib.ZSTD_decompress(buf, int(size), data, len(data))
    if lib.ZSTD_isError(got):
        raise ValueError("zstd frame could not be inflated")
    return buf.raw[:got]


def zstd_compress(data: bytes, level: int = 3) -> bytes:
    lib = _libzstd()
    cap = lib.ZSTD_compressBound(len(data))
    buf = ctypes.create_string_buffer(cap)
    got = lib.ZSTD_compress(buf, cap, data, len(data), level)
    if lib.ZSTD_isError(got):
        raise ValueError("zstd compression failed")
    return buf.raw[:got]


# --------------------------------------------------------------------------- decode
def _unquote(data: bytes, pos: int):
    """``std::quoted(s, '"', '\\\\')`` extraction starting at the opening quote."""
    assert data[pos : pos + 1] == b'"', "record does not start with a quote"
    pos += 1
    out = bytearray()
    n = len(data)
    while pos < n:
        c = data[pos]
        if c == 0x5C:  # backslash: next byte is literal
            out.append(data[pos + 1])
            pos += 2
        elif c == 0x22:  # closing quote
            return bytes(out), pos + 1
        else:
            out.append(c)
            pos += 1
    raise ValueError("unterminated body record")


def loads(data: bytes) -> Snapshot:
    compressed = data[:4] == ZSTD_MAGIC
    if compressed:
        data = zstd_decompress(data)

    sep = data.find(b"- - -")
    if sep < 0:
        raise ValueError("snapshot has no '- - -' separator")
    header: Dict[str, str] = {}
    for line in data[:sep].decode("latin-1").splitlines():
        if "=" in line:
            k, v = line.split("=", 1)
            header[k.strip()] = v.strip().strip('"')

    version = header.get("MODEL_VERSION", "0.0.1")
    vt = _version_tuple(version)
    if vt > _version_tuple(MODEL_VERSION):  # World_SaveLoad.cpp:102-106
        raise ValueError("Unsupported snapshot version:" + version)
    friction = float(np.float32(float(header["drag"])))  # stof
    interact = bool(float(header["interactions"]))
    gravity_mode = GRAVITY_HYPERBOLIC  # Properties::defaults, World.cpp:42-50
    gravity = PHYS_G
    if vt >= (0, 1, 0):
        gravity_mode = int(header["gravity_mode"])
    if vt >= (0, 1, 2):  # 0.1.1 saved G cast to unsigned; ignored (World_SaveLoad.cpp:137)
        gravity = float(np.float32(float(header["gravity_strength"])))
    n = int(header["objects"])

    pos = data.index(b"\n", sep) + 1
    raws = []
    for i in range(n):
        colon = data.index(b":", pos)
        ndx = int(data[pos:colon].strip())
        assert ndx == i, "body index out of sequence"
        q = data.index(b'"', colon)
        raw, pos = _unquote(data, q)
        raws.append(raw)
        while pos < len(data) and data[pos] in b" \r\n\t":
            pos += 1
    size = len(raws[0]) if raws else BODY_F32.itemsize
    if any(len(r) != size for r in raws):
        raise ValueError("body records of unequal size")
    if size == BODY_F32.itemsize:
        dt = BODY_F32
    elif size == BODY_F64.itemsize:
        dt = BODY_F64
    else:  # Entity_SaveLoad.cpp:66-69
        raise ValueError("Failed to load object! Bytes expected: 60 or 96, found: %d" % size)
    bodies = np.frombuffer(b"".join(raws), dtype=dt).copy() if raws else np.zeros(0, dt)
    return Snapshot(version, friction, interact, gravity_mode, gravity, bodies, header, compressed)


def load(path: str) -> Snapshot:
    with open(path, "rb") as f:
        return loads(f.read())


# --------------------------------------------------------------------------- encode
def _fmt_ostream(x: float) -> str:
    """``std::ostream << float/double`` with default flags == ``%g`` (6 significant digits)."""
    return "%g" % x


def dumps(s: Snapshot, version: Optional[str] = None, compress: bool = False) -> bytes:
    version = version or MODEL_VERSION
    vt = _version_tuple(version)
    out = bytearray()
    out += b"MODEL_VERSION = " + version.encode() + b"\n"
    out += b"drag = " + _fmt_ostream(np.float32(s.friction)).encode() + b"\n"
    out += b"interactions = " + (b"1" if s.interact_n2n else b"0") + b"\n"
    if vt >= (0, 1, 0):
        out += b"gravity_mode = %d\n" % int(s.gravity_mode)
    if vt >= (0, 1, 2):
        g = np.float32(s.gravity) if s.bodies.dtype == BODY_F32 else s.gravity
        out += b"gravity_strength = " + _fmt_ostream(g).encode() + b"\n"
    out += b"objects = %d\n" % s.n
    out += b"- - -\n"
    raw = s.bodies.tobytes()
    size = s.bodies.dtype.itemsize
    for i in range(s.n):
        rec = raw[i * size : (i + 1) * size].replace(b"\\", b"\\\\").replace(b'"', b'\\"')
        out += b"%d : \"" % i + rec + b"\"\n"
    data = bytes(out)
    return zstd_compress(data) if compress else data


def save(path: str, s: Snapshot, version: Optional[str] = None, compress: bool = False) -> None:
    with open(path, "wb") as f:
        f.write(dumps(s, version, compress))
