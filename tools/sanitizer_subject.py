#!/usr/bin/env python3
"""Subject for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel of the library on tiny worlds.

    compute-sanitizer --tool racecheck python tools/sanitizer_subject.py            # one GPU
    compute-sanitizer --target-processes all --tool memcheck python tools/sanitizer_subject.py --ranks 2

Worlds: KAT-1 (1010 bodies = 4 tiles, the reference's tc-smoke start state) and a ragged 3-tile world (600 bodies, 20 % of
them mortal: terminate + sweep + compaction), FAST and EXACT, all-pairs and player-only, half and full matrix, realistic
law + repulsion, the emitters, the render read-back, table edits.  A 40 000-body cloud adds the Hilbert order kernels and
the sort-ahead stream (two frames).  Results are not checked here (the parity tests do that): the sanitizer is.
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def subject(rank, nranks, nccl_id, big):
    from out_of_nothing_b200 import capi, snapshot, synth
    from out_of_nothing_b200.world import World
    kat1 = snapshot.load(os.path.join(ROOT, "tests", "golden", "kat1_start.state"))
    ragged, cfg = synth.make("c3_dense_13k", n=600)
    ragged["lifetime"][1:120] = 0.05          # expire in the second frame, in runs (the sweep's skip-after-erase quirk)
    dt = 0.033
    props = dict(friction=0.01, gravity_mode=1, gravity=kat1.gravity, interact_n2n=1)
    for mode in (capi.MATH_FAST, capi.MATH_EXACT):
        for bodies, extra in ((kat1.bodies, {}), (ragged, {}), (ragged, dict(loop_mode=1)), (kat1.bodies, dict(interact_n2n=0)),
                              (ragged, dict(gravity_mode=2, repulsion_stiffness=1e-3))):
            w = World(rank, rank, nranks, nccl_id)
            w.set_props(**{**props, **extra})
            w.math_mode = mode
            w.debug_capture_pairs(1 << 14)
            w.upload(bodies)
            w.step(dt, 3)
            w.update_intrinsic(dt); w.update_pairwise(dt); w.update_global(dt); w.sweep_expired(0)
            w.download(); w.download_render(); w.drain_events(); w.entity_count()
            w.debug_fetch_pairs(0); w.debug_fetch_pairs(1)
            if nranks == 1:
                cfg_e = w.emitter_config(particle_lifetime=0.05, create_mass=0, particle_mass_min=1e20, particle_mass_max=2e20)
                w.emit_particles(0, 7, cfg_e, seed=3)
                w.chemtrail_burst(0, 5, cfg_e, seed=4)
                w.spawn(0, 3, seed=5)
                w.step(dt, 2)
                b = w.entity(1); w.set_entity(1, b); w.remove_body(2); w.add_body(b)
                w.step(dt, 1)
                w.download()
            w.close()
    if big:
        bodies, cfg = synth.make("c4_cloud_100k", n=40000)
        w = World(rank, rank, nranks, nccl_id)
        w.set_props(**cfg["props"])
        w.upload(bodies)
        w.step(cfg["dt"], 2)
        w.download()
        w.close()
    print("sanitizer subject: rank %d of %d done" % (rank, nranks), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--big", action="store_true", help="add the 40 000-body cloud (Hilbert order kernels, sort-ahead stream)")
    args = ap.parse_args()
    if args.ranks == 1:
        subject(0, 1, None, args.big)
        return 0
    from out_of_nothing_b200 import capi
    nccl_id = capi.nccl_unique_id()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=subject, args=(r, args.ranks, nccl_id, args.big)) for r in range(args.ranks)]
    for p in procs:
        p.start()
    rc = 0
    for p in procs:
        p.join(1500)
        rc |= p.exitcode or 0
    return rc


if __name__ == "__main__":
    sys.exit(main())
