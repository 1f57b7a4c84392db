"""Single-process multi-GPU handle (oon_create_multi): one host thread, N GPUs, peer-access pointers instead of CUDA IPC,
in-process event/peer-copy collectives instead of NCCL — what the reference's single-process app or the headless driver can
call (main.cpp:88-91, World.cpp:196-198).  Everything a single-GPU handle does must work on it and give the same bits.

Needs >= 2 GPUs; skipped otherwise.  `gpurun --gpus 2 -- python -m pytest tests/test_gpu_group.py -m gpu`.
"""
import numpy as np
import pytest

from helpers import bodies_equal_except_color, load_golden, make_oracle, pair_list, props_of
from out_of_nothing_b200 import capi, synth

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

FAST, EXACT = capi.MATH_FAST, capi.MATH_EXACT
DT = 0.033


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def ndev():
    n = min(gpu_count(), 8)
    return 1 << (n.bit_length() - 1) if n else 0


needs2 = pytest.mark.skipif(gpu_count() < 2, reason="needs 2 GPUs")


def world(bodies, props, mode, devices=None, capture=0):
    from out_of_nothing_b200.world import World
    w = World(0) if devices is None else World(devices=devices)
    w.set_props(**{k: (int(v) if isinstance(v, bool) else v) for k, v in props.items()})
    w.math_mode = mode
    if capture:
        w.debug_capture_pairs(capture)
    w.upload(bodies)
    return w


def pair(bodies, props, mode, capture=0):
    return world(bodies, props, mode, None, capture), world(bodies, props, mode, list(range(ndev())), capture)


@needs2
@pytest.mark.parametrize("mode", [FAST, EXACT])
@pytest.mark.parametrize("name", ["kat1", "cloud20k", "cloud50k", "player_only"])
def test_group_is_bit_identical_to_one_gpu(mode, name):
    if name == "kat1":
        s = load_golden("kat1_start.state")
        bodies, props, steps = s.bodies, props_of(s, friction=0.01, interact_n2n=True), 20
    elif name == "player_only":
        s = load_golden("kat2_start.state")
        bodies = s.bodies.copy(); bodies["gravity_immunity"][0] = 0
        props, steps = props_of(s, friction=0.01), 10
    else:
        n = 20000 if name == "cloud20k" else 50000
        if n == 50000 and mode == EXACT:
            pytest.skip("EXACT at 50k adds nothing over 20k")
        bodies, cfg = synth.make("c4_cloud_100k", n=n)
        props, steps = cfg["props"], 6
    small = len(bodies) < 5000                                # (a 20k-body cloud has millions of colliding pair-frames: counts only)
    one, grp = pair(bodies, props, mode, capture=(1 << 20) if small else 0)
    assert grp.device_count() == ndev() and one.device_count() == 1
    assert grp.exchange_kind() == "peer_stores"
    one.step(DT, steps); grp.step(DT, steps)
    a, b = one.download(), grp.download()
    assert a.tobytes() == b.tobytes()
    xa, ca = one.download_render_rgb(); xb, cb = grp.download_render_rgb()
    assert xa.tobytes() == xb.tobytes() and ca.tobytes() == cb.tobytes()
    assert (xa[:, :2] == a["p"]).all() and (xa[:, 2] == a["r"]).all() and (ca == a["color"]).all()
    if small:
        assert pair_list(one.debug_fetch_pairs(0)) == pair_list(grp.debug_fetch_pairs(0))
        assert pair_list(one.debug_fetch_pairs(1)) == pair_list(grp.debug_fetch_pairs(1))
    ea, eb = one.drain_events(), grp.drain_events()
    for k in ("steps", "collisions", "touches", "player_touches", "terminated", "removed"):
        assert ea[k] == eb[k], k
    # the phase-by-phase API and the shard-local calls (to its caller the group owns the whole table)
    for w in (one, grp):
        w.update_intrinsic(DT); w.update_pairwise(DT); w.update_global(DT); w.sweep_expired(0)
    assert one.download().tobytes() == grp.download().tobytes()
    first, block = grp.download_local()
    assert first == 0 and block.tobytes() == one.download().tobytes()
    grp.upload_local(block, len(block)); one.upload(block)
    one.step(DT, 2); grp.step(DT, 2)
    assert one.download().tobytes() == grp.download().tobytes()
    one.close(); grp.close()


def _boundary_runs():
    bodies, cfg = synth.make("c4_cloud_100k", n=3000)
    bodies = bodies.copy()
    life = bodies["lifetime"]
    for lo, hi, frame in ((2040, 2056, 1), (2047, 2049, 2), (2030, 2048, 3), (2048, 2070, 4), (1, 40, 2), (2990, 3000, 3),
                          (2046, 2051, 5), (1000, 1003, 1), (2044, 2053, 7), (760, 775, 2), (1530, 1540, 3)):
        life[lo:hi] = np.float32(DT * frame - 0.001)
    rng = np.random.default_rng(3)
    pick = rng.random(3000) < 0.2
    pick[0] = False
    life[pick] = rng.uniform(0.01, 0.4, pick.sum()).astype(np.float32)
    return bodies, cfg["props"]


@needs2
def test_group_sweep_equals_one_gpu_and_the_oracle():
    """Runs of expired bodies across the shards' block boundaries: same removal sequence (indices at the time of removal),
    same survivors, EXACT mode bit-identical to one GPU and to the oracle."""
    bodies, props = _boundary_runs()
    one, grp = pair(bodies, props, EXACT)
    o = make_oracle(bodies, props, trace=True)
    for frame in range(14):
        o.trace_clear()
        o.step(DT, 1); one.step(DT, 1); grp.step(DT, 1)
        expired = o.trace_fetch()["expired"].tolist()
        assert one.debug_fetch_removed().tolist() == expired, frame
        assert grp.debug_fetch_removed().tolist() == expired, frame
        assert grp.entity_count() == o.n
    ok, field = bodies_equal_except_color(grp.download(), o.bodies())
    assert ok, field
    assert grp.download().tobytes() == one.download().tobytes()
    assert o.n < 3000 - 200
    one.close(); grp.close()


@needs2
@pytest.mark.parametrize("mode", [FAST, EXACT])
def test_group_table_edits_and_emitters_follow_one_gpu(mode):
    """add / remove / set / get, the three body-creating loops and a thrusting player whose exhaust expires, on a group: the
    edits are routed to the owning shard.  EXACT: bit-identical to one GPU and to the oracle; FAST: every decision, T,
    lifetimes and counts identical (after edits the shards' blocks are ragged, so FP32 sums round differently)."""
    s = load_golden("kat1_start.state")
    bodies = s.bodies.copy()
    bodies["thrust"][0] = [3e32, 0, 1e32, 0]                       # the player burns: thrust kick every frame
    props = props_of(s, friction=0.01, interact_n2n=True)
    one, grp = pair(bodies, props, mode)
    o = make_oracle(bodies, props)
    worlds = (one, grp)
    cfg = one.emitter_config(particle_lifetime=0.12, create_mass=0, particle_mass_min=1e20, particle_mass_max=3e20,
                             eject_velocity=(0.0, -4e8), velocity_divergence=2.0)
    extra = synth.random_cloud(1501, seed=11)[1:]
    for frame in range(24):
        if frame == 3:
            for w in worlds:
                w.add_bodies(extra[:300])
            o.set_bodies(np.concatenate([o.bodies(), extra[:300]]))
        if frame == 5:
            for w in worlds:
                w.remove_body(40); w.remove_body(900); w.remove_body(w.entity_count() - 1)
            b = o.bodies(); b = np.delete(b, 40); b = np.delete(b, 900); b = b[:-1]
            o.set_bodies(b)
        if frame == 7:
            e = one.entity(1200).copy()
            assert e.tobytes() == grp.entity(1200).tobytes()
            e["v"] = [1e7, -2e7]; e["lifetime"] = 0.2
            for w in worlds:
                w.set_entity(1200, e)
            b = o.bodies(); b[1200] = e; o.set_bodies(b)
        if frame == 9:
            for w in worlds:
                w.add_bodies(extra[300:])                              # more than the tail shard has room for: re-layout
                w.spawn(5, 4, seed=77)
                w.chemtrail_burst(1100, 6, cfg, seed=78)
            o.set_bodies(np.concatenate([o.bodies(), extra[300:]]))
            o.spawn(5, 4, seed=77); o.chemtrail_burst(1100, 6, cfg, seed=78)
        got = [w.emit_particles(0, 5, cfg, seed=1000 + frame) for w in worlds]
        assert got[0] == got[1] == o.emit_particles(0, 5, cfg, seed=1000 + frame)
        if mode == FAST:                                              # teacher-forced on the oracle's state
            state = o.bodies()
            for w in worlds:
                w.upload(state)
        o.step(DT, 1)
        for w in worlds:
            w.step(DT, 1)
        ref = o.bodies()
        for w in worlds:
            assert w.entity_count() == o.n, frame
            gotb = w.download()
            if mode == EXACT:
                ok, field = bodies_equal_except_color(gotb, ref)
                assert ok, (frame, field)
            else:
                for f in ("T", "lifetime", "r", "mass", "density"):
                    assert (gotb[f] == ref[f]).all(), (frame, f)
    assert o.n > len(bodies) + 1400                                    # the world did grow, and exhaust did expire
    assert one.drain_events()["removed"] == grp.drain_events()["removed"] > 50
    one.close(); grp.close()


@needs2
def test_group_timing_hooks_and_errors():
    bodies, cfg = synth.make("c4_cloud_100k", n=40000)
    one, grp = pair(bodies, cfg["props"], FAST)
    for w in (one, grp):
        w.step(cfg["dt"], 3)
    ms1 = one.bench_frames(cfg["dt"], 4, 64 << 20); msg = grp.bench_frames(cfg["dt"], 4, 64 << 20)
    assert 0 < msg < ms1                                              # more GPUs, less time per frame
    _, prof, launches = grp.bench_frames_profiled(cfg["dt"], 2, 64 << 20)
    assert prof["interact"] > 0 and launches["interact"] == 2
    grp.timer_start(); grp.step(cfg["dt"], 2); assert grp.timer_stop() > 0
    one.step(cfg["dt"], 4)                                            # the same number of frames on both
    assert one.download().tobytes() == grp.download().tobytes()
    assert grp.launch_count() > one.launch_count()
    with pytest.raises(capi.OonError):
        grp.upload_local(bodies[:10], len(bodies))                    # a group owns the whole table
    with pytest.raises(capi.OonError):
        from out_of_nothing_b200.world import World
        World(devices=[0, 0])
    one.close(); grp.close()


@needs2
def test_group_large_world_with_expiring_bodies_frame_by_frame():
    """40 000 bodies, 5 % of them expiring within a few frames: the curve order is rebuilt after every sweep on every shard.  From
    the same uploaded state a frame of the group equals a frame of one GPU bit for bit (FAST), sweep included."""
    bodies, cfg = synth.make("c4_cloud_100k", n=40000)
    bodies = bodies.copy()
    rng = np.random.default_rng(11)
    pick = rng.random(40000) < 0.05
    pick[0] = False
    bodies["lifetime"][pick] = rng.uniform(0.01, 0.2, pick.sum()).astype(np.float32)
    one, grp = pair(bodies, cfg["props"], FAST)
    state = bodies
    for frame in range(5):
        if frame:
            one.upload(state); grp.upload(state)
        one.step(cfg["dt"], 1); grp.step(cfg["dt"], 1)
        assert one.debug_fetch_removed().tolist() == grp.debug_fetch_removed().tolist(), frame
        a, b = one.download(), grp.download()
        assert a.tobytes() == b.tobytes(), frame
        state = a
    assert len(state) < 40000 - 500
    # ... and free-running from there: same removals and lifetimes (the sums round differently once the blocks are ragged)
    one.step(cfg["dt"], 4); grp.step(cfg["dt"], 4)
    a, b = one.download(), grp.download()
    assert len(a) == len(b) and a["lifetime"].tobytes() == b["lifetime"].tobytes()
    one.close(); grp.close()
