"""The C++ host mirror (out_of_nothing_b200/host): snapshot codec and the headless driver's command line.

CPU part: the codec round-trips the reference's fixtures byte for byte and the driver fails loudly without a device.
GPU part: the reference's tc-smoke regression run (test/regression/tc-smoke.cmd:30-43) replayed through oon_headless.
"""
import os
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN, TC_SMOKE, bodies_equal_except_color, load_golden, make_oracle, props_of
from out_of_nothing_b200 import snapshot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "out_of_nothing_b200", "host")
EXE = os.path.join(HOST, "oon_headless")


@pytest.fixture(scope="module")
def exe():
    subprocess.run(["make", "-C", os.path.join(ROOT, "out_of_nothing_b200", "csrc"), "-s"], check=True)
    subprocess.run(["make", "-C", HOST, "-s"], check=True)
    return EXE


def run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True, text=True)


@pytest.mark.parametrize("name", ["kat1_start.state", "kat1_end.state", "kat2_start.state", "uncompressed_0.0.1.sav"])
def test_cpp_codec_reproduces_uncompressed_fixtures(exe, tmp_path, name):
    out = tmp_path / "out.state"
    r = run(exe, "--convert-only", "--session=" + os.path.join(GOLDEN, name), "--session-save-as=" + str(out), "--no-save-compressed")
    assert r.returncode == 0, r.stderr
    assert out.read_bytes() == open(os.path.join(GOLDEN, name), "rb").read()


@pytest.mark.parametrize("name", ["kat3_start.state", "multi_cluster_640.save", "snapshot_4_14k.save", "one_body_realg_double.save"])
def test_cpp_codec_zstd_and_double_records(exe, tmp_path, name):
    out = tmp_path / "out.state"
    r = run(exe, "--convert-only", "--session=" + os.path.join(GOLDEN, name), "--session-save-as=" + str(out))
    assert r.returncode == 0, r.stderr
    a, b = snapshot.load(os.path.join(GOLDEN, name)), snapshot.load(str(out))
    assert b.compressed and b.n == a.n
    assert b.bodies.tobytes() == a.as_f32().tobytes()
    assert (b.friction, b.interact_n2n, b.gravity_mode) == (a.friction, a.interact_n2n, a.gravity_mode)


def test_cpp_loader_rejects_newer_versions(exe, tmp_path):
    raw = open(os.path.join(GOLDEN, "kat1_start.state"), "rb").read().replace(b"MODEL_VERSION = 0.1.2", b"MODEL_VERSION = 7.0.0")
    bad = tmp_path / "bad.state"; bad.write_bytes(raw)
    r = run(exe, "--convert-only", "--session=" + str(bad), "--session-save-as=" + str(tmp_path / "x"))
    assert r.returncode != 0 and "Unsupported snapshot version" in r.stderr


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_driver_has_no_cpu_path(exe):
    r = run(exe, "--session=" + os.path.join(GOLDEN, "kat1_start.state"), "--loop-cap=1")
    assert r.returncode != 0 and "no CPU path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("math", ["exact", "fast"])
def test_tc_smoke_through_the_headless_driver(exe, tmp_path, math):
    end = tmp_path / "END.state.tmp"
    r = run(exe, "--headless", "--cfg=test/default.cfg", "--interact", "--friction=0.01", "--zoom-adjust=0.2", "--fixed-dt=0.033",
            "--fps-limit=0", "--loop-cap=20", "--exit-on-finish", "--session=" + os.path.join(GOLDEN, "kat1_start.state"),
            "--session-save-as=" + str(end), "--no-save-compressed", "--log-level=D", "--math=" + math, "--report")
    assert r.returncode == 0, r.stderr
    got = snapshot.load(str(end))
    ref_end = load_golden("kat1_end.state")
    assert got.n == 1010 and got.version == "0.1.5" and not got.compressed
    assert (got.bodies["T"] == ref_end.bodies["T"]).all()              # the reference's own END state: every decision matches
    for f in ("r", "mass", "density", "lifetime", "gravity_immunity", "free_color", "thrust"):
        assert got.bodies[f].tobytes() == ref_end.bodies[f].tobytes(), f
    start = load_golden("kat1_start.state")
    o = make_oracle(start.bodies, props_of(start, friction=TC_SMOKE["friction"], interact_n2n=True)); o.step(TC_SMOKE["dt"], 20)
    if math == "exact":
        ok, field = bodies_equal_except_color(got.bodies, o.bodies())
        assert ok, field
    assert '"collisions": 3590' in r.stdout and '"touches": 322' in r.stdout


@pytest.mark.gpu
def test_headless_bodies_flag_and_realg(exe, tmp_path):
    end = tmp_path / "e.state"
    r = run(exe, "--headless", "--bodies=500", "--interact", "--g-mode=R", "--friction=0", "--fixed-dt=0.001", "--loop-cap=50",
            "--session-save-as=" + str(end), "--report")                # test-headless.cmd / test-realg.cmd style
    assert r.returncode == 0, r.stderr
    s = snapshot.load(str(end))
    assert s.n == 501 and s.gravity_mode == 2 and s.compressed and s.interact_n2n
    assert np.isfinite(s.bodies["p"]).all()
