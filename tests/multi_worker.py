"""Worker of the multi-GPU parity test: one process per GPU, one shard each (launched by test_gpu_multi.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rank, nranks, nccl_id, bodies_path, out_path, mode, steps, dt, props, p2p=True, split_steps=False, last_frame_fast=False):
    os.environ["OON_P2P"] = "1" if p2p else "0"
    from out_of_nothing_b200 import capi
    from out_of_nothing_b200.world import World
    bodies = np.load(bodies_path)
    w = World(rank, rank, nranks, nccl_id)
    w.set_props(**props)
    w.math_mode = mode
    w.upload(bodies)
    if split_steps:                      # the three virtuals one at a time + a re-upload of the same table in between
        for k in range(steps):
            w.update_intrinsic(dt); w.update_pairwise(dt); w.update_global(dt)
            if k == steps // 2:
                w.upload(w.download())
            if k == steps // 2 + 1 and not (bodies["lifetime"] > 0).any():      # the same through the shard-local calls: own block only
                n_total = w.entity_count()
                first, block = w.download_local()
                assert (first, len(block)) == w.local_range(n_total)
                w.upload_local(block, n_total)
    elif last_frame_fast:                # `steps` frames in the given mode, then one FAST frame from that (shared) state
        w.step(dt, steps)
        w.math_mode = capi.MATH_FAST
        w.step(dt, 1)
    else:
        w.step(dt, steps)
    out = w.download()
    first, block = w.download_local()                      # shard-local read-back == this rank's slice of the gathered table
    assert block.tobytes() == out[first:first + len(block)].tobytes(), "download_local differs from download on rank %d" % rank
    xyr = w.download_render()
    ev = w.drain_events()
    np.save(out_path % rank, out)
    np.save((out_path % rank) + ".xyr.npy", xyr)
    np.save((out_path % rank) + ".ev.npy", np.array([ev["collisions"], ev["touches"], ev["terminated"], ev["removed"]], np.int64))
    with open((out_path % rank) + ".kind", "w") as f:
        f.write(w.exchange_kind())
    w.close()


def run_edits(rank, nranks, nccl_id, out_path, mode):
    """Table edits and the body-creating loops on a sharded world, every rank making the same calls (collective, like upload):
    a thrusting player whose exhaust expires, appends that overflow the tail shard, removals, a GUI-style edit."""
    from out_of_nothing_b200 import snapshot, synth
    from out_of_nothing_b200.world import World
    s = snapshot.load(os.path.join(ROOT, "tests", "golden", "kat1_start.state"))
    bodies = s.bodies.copy()
    bodies["thrust"][0] = [3e32, 0, 1e32, 0]
    w = World(rank, rank, nranks, nccl_id)
    w.set_props(friction=0.01, gravity_mode=s.gravity_mode, gravity=s.gravity, interact_n2n=1)
    w.math_mode = mode
    w.upload(bodies)
    cfg = w.emitter_config(particle_lifetime=0.12, create_mass=0, particle_mass_min=1e20, particle_mass_max=3e20,
                           eject_velocity=(0.0, -4e8), velocity_divergence=2.0)
    extra = synth.random_cloud(1501, seed=11)[1:]
    got = []
    for frame in range(20):
        if frame == 3:
            w.add_bodies(extra[:300])
        if frame == 5:
            w.remove_body(40); w.remove_body(900); w.remove_body(w.entity_count() - 1)
        if frame == 7:
            e = w.entity(1200).copy()
            e["v"] = [1e7, -2e7]; e["lifetime"] = 0.2
            w.set_entity(1200, e)
        if frame == 9:
            w.add_bodies(extra[300:])
            w.spawn(5, 4, seed=77)
            w.chemtrail_burst(1100, 6, cfg, seed=78)
        got.append(w.emit_particles(0, 5, cfg, seed=1000 + frame))
        w.step(0.033, 1)
    out = w.download()
    ev = w.drain_events()
    np.save(out_path % rank, out)
    np.save((out_path % rank) + ".ev.npy", np.array([ev["collisions"], ev["touches"], ev["terminated"], ev["removed"], sum(got)], np.int64))
    w.close()
