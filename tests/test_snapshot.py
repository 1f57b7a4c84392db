"""Snapshot codec against every fixture the reference ships (World_SaveLoad.cpp, Entity_SaveLoad.cpp)."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN
from out_of_nothing_b200 import snapshot as S

MANIFEST = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))

EXPECT = {  # name -> (N, record bytes, version, n2n, gravity_mode)
    "kat1_start.state": (1010, 60, "0.1.2", True, 1),
    "kat1_end.state": (1010, 60, "0.1.4", True, 1),
    "kat2_start.state": (1010, 60, "0.1.2", False, 1),
    "kat3_start.state": (501, 60, "0.0.1", True, 1),
    "kat3_end.state": (501, 60, "0.1.4", True, 1),
    "realg_test.state": (40, 60, "0.1.3", True, 2),
    "multi_cluster_640.save": (691, 60, "0.1.1", True, 1),
    "two_bodies_realg.state": (3, 60, "0.1.3", True, 3),
    "earth_moon.state": (102, 60, "0.1.2", True, 0),
    "one_body_realg_double.save": (1, 96, "0.1.3", True, 2),
    "uncompressed_0.0.1.sav": (101, 60, "0.0.1", False, 1),
    "snapshot_4_14k.save": (14068, 60, "0.0.1", False, 1),
}


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_every_fixture_decodes(name):
    s = S.load(os.path.join(GOLDEN, name))
    assert s.n == int(s.header["objects"])
    assert s.bodies.dtype.itemsize in (60, 96)
    if name in EXPECT:
        n, size, ver, n2n, gm = EXPECT[name]
        assert (s.n, s.bodies.dtype.itemsize, s.version, s.interact_n2n, s.gravity_mode) == (n, size, ver, n2n, gm)
    assert np.isfinite(s.bodies["mass"]).all() and (s.bodies["mass"] > 0).all()


@pytest.mark.parametrize("name,version", [("kat1_start.state", "0.1.2"), ("kat1_end.state", "0.1.4"),
                                          ("kat2_start.state", "0.1.2"), ("uncompressed_0.0.1.sav", "0.0.1")])
def test_writer_reproduces_uncompressed_fixtures_byte_for_byte(name, version):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    s = S.loads(raw)
    assert not s.compressed
    assert S.dumps(s, version=version) == raw


def test_zstd_roundtrip_and_version_gates():
    s = S.load(os.path.join(GOLDEN, "kat3_start.state"))
    assert s.compressed and s.version == "0.0.1"
    assert s.gravity_mode == 1 and abs(s.gravity - 6.6743e-11) < 1e-20   # defaults: fields absent before 0.1.0 / 0.1.2
    again = S.loads(S.dumps(s, compress=True))
    assert again.compressed and again.bodies.tobytes() == s.bodies.tobytes()
    assert again.version == S.MODEL_VERSION


def test_quote_and_backslash_escaping():
    b = np.zeros(3, S.BODY_F32)
    b["color"] = [0x22225C5C, 0x5C225C22, 0x22]     # bytes that need escaping: '"' = 0x22, '\\' = 0x5C
    b["mass"] = 1.0
    snap = S.Snapshot("0.1.5", 0.03, True, 1, 6.6743e-11, b)
    assert S.loads(S.dumps(snap)).bodies.tobytes() == b.tobytes()


def test_loader_errors_like_the_reference():
    raw = open(os.path.join(GOLDEN, "kat1_start.state"), "rb").read()
    with pytest.raises(ValueError, match="Unsupported snapshot version"):      # World_SaveLoad.cpp:102-106
        S.loads(raw.replace(b"MODEL_VERSION = 0.1.2", b"MODEL_VERSION = 9.0.0"))
    with pytest.raises(ValueError):                                            # no separator
        S.loads(raw.replace(b"- - -", b"-----"))
    with pytest.raises((ValueError, AssertionError)):                          # truncated record -> size mismatch
        S.loads(raw[: len(raw) // 2] + b'"\n')


def test_double_build_record_narrows():
    s = S.load(os.path.join(GOLDEN, "one_body_realg_double.save"))
    assert s.bodies.dtype == S.BODY_F64
    f = s.as_f32()
    assert f.dtype == S.BODY_F32 and f["mass"][0] == np.float32(s.bodies["mass"][0])


def test_golden_manifest_matches_files():
    import hashlib
    for name, meta in MANIFEST.items():
        data = open(os.path.join(GOLDEN, name), "rb").read()
        assert hashlib.sha256(data).hexdigest() == meta["sha256"], name
