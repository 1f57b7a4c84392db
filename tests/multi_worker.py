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
