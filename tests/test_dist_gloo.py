"""world_size-2 test of the sharding scheme on CPU (gloo): target blocks per rank + one all-gather of the source
records per frame reproduce the reference's frame.  The per-shard arithmetic here is a float64 numpy restatement
(test code); the device kernels implement the same data flow with NCCL (tests/test_gpu_multi.py on >= 2 GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import kick_f64_sample, load_golden, make_oracle, props_of
from out_of_nothing_b200 import sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        s = load_golden("kat1_start.state")
        props = props_of(s, friction=0.01, interact_n2n=True)
        bodies, n, dt = s.bodies, s.n, np.float32(0.033)
        # an opaque 128-byte id travels from rank 0 to everyone, like the NCCL unique id in bench.py
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(idt, 0)
        assert idt.tolist() == list(range(128))

        block = sharding.block_size(n, world_size)
        lo, hi = sharding.shard_range(n, rank, world_size)
        # exchange record per slot: x, y, m, r (dead/padding slots: m = 0, r < 0) — one all-gather per frame
        mine = torch.zeros(block, 4, dtype=torch.float64)
        mine[:, 3] = -1.0
        mine[: hi - lo, 0:2] = torch.from_numpy(bodies["p"][lo:hi].astype(np.float64))
        mine[: hi - lo, 2] = torch.from_numpy(bodies["mass"][lo:hi].astype(np.float64))
        mine[: hi - lo, 3] = torch.from_numpy(bodies["r"][lo:hi].astype(np.float64))
        gathered = [torch.zeros_like(mine) for _ in range(world_size)]
        dist.all_gather(gathered, mine)
        allsrc = torch.cat(gathered).numpy()
        live = allsrc[:, 3] >= 0
        assert live.sum() == n
        src = np.zeros(n, bodies.dtype)
        src["p"], src["mass"], src["r"] = allsrc[live, 0:2], allsrc[live, 2], allsrc[live, 3]
        src["gravity_immunity"] = bodies["gravity_immunity"]            # private per-target flag, only the owner's rows are used
        # own targets only: kick from every source, then drag + drift (World.cpp:447-462)
        kick = kick_f64_sample(src, props, dt, np.arange(lo, hi))
        v = bodies["v"][lo:hi].astype(np.float64) + kick
        v -= v * float(np.float32(props["friction"])) * float(dt)
        p = bodies["p"][lo:hi].astype(np.float64) + v * float(dt)
        np.save(os.path.join(out_dir, "pv_%d.npy" % rank), np.concatenate([p, v], axis=1))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_frame_matches_the_oracle(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.concatenate([np.load(tmp_path / ("pv_%d.npy" % r)) for r in range(2)])
    s = load_golden("kat1_start.state")
    o = make_oracle(s.bodies, props_of(s, friction=0.01, interact_n2n=True), double=True)
    o.step(0.033, 1)
    ref = o.pv64()
    assert got.shape == ref.shape
    scale = np.abs(ref).max(axis=0)
    assert (np.abs(got - ref) / scale).max() < 1e-9


def test_shard_geometry_is_canonical():
    """1, 2, 4 and 8 ranks cut the world at the same tile-aligned boundaries (8 canonical blocks)."""
    for n in (1010, 10703, 100000, 1000000):
        cuts = {g: {sharding.shard_range(n, r, g)[0] for r in range(g)} - {n} for g in (1, 2, 4, 8)}
        assert cuts[1] <= cuts[2] <= cuts[4] <= cuts[8]
        assert all(c % sharding.TS == 0 for c in cuts[8])


# ------------------------------------------------------------------------------------------------ sharded sweep
def _sweep_worker(rank, world_size, port, out_dir):
    """The expired-body sweep of a sharded world: local scan, ONE all-gather of the ranks' run triples, local compaction,
    then an all-gather of the survivor counts — the data flow of do_sweep()/sync_count() in csrc/oon_capi.cu, with gloo."""
    from test_sweep_model import UNLIMITED, scan_sweep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        rng = np.random.default_rng(7)                               # same table on every rank, like oon_upload
        n = 3000
        life = np.full(n, UNLIMITED, np.float32)
        life[rng.random(n) < 0.3] = 0
        life[2040:2060] = 0                                          # a run across the boundary of the 2-rank blocks (2048)
        life[0] = UNLIMITED
        lo, hi = sharding.shard_range(n, rank, world_size)
        mine = life[lo:hi]
        first_tested = 1 if rank == 0 else 0                         # the player (index 0) lives on rank 0
        _, triple = scan_sweep(mine, first_tested)
        t = torch.tensor([triple[0], triple[1], int(triple[2]), 0], dtype=torch.int32)
        gathered = [torch.zeros_like(t) for _ in range(world_size)]
        dist.all_gather(gathered, t)                                 # 16 bytes per rank
        carry = 0
        for h in range(rank - 1, -1, -1):
            carry += int(gathered[h][1])
            if not int(gathered[h][2]):
                break
        keep, _ = scan_sweep(mine, first_tested, carry)
        survivors = (lo + np.nonzero(keep)[0]).astype(np.int64)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
        dist.all_gather(counts, torch.tensor([len(survivors)], dtype=torch.int64))
        prefix = int(sum(int(c) for c in counts[:rank]))             # global index of this rank's first body after the sweep
        np.save(os.path.join(out_dir, "surv_%d.npy" % rank), survivors)
        np.save(os.path.join(out_dir, "prefix_%d.npy" % rank), np.array([prefix, int(sum(int(c) for c in counts))]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_sweep_matches_the_literal_erase_loop(tmp_path):
    from test_sweep_model import UNLIMITED, literal_sweep
    port = _free_port()
    mp.spawn(_sweep_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(7)
    n = 3000
    life = np.full(n, UNLIMITED, np.float32)
    life[rng.random(n) < 0.3] = 0
    life[2040:2060] = 0
    life[0] = UNLIMITED
    _, alive = literal_sweep(life)
    got = np.concatenate([np.load(tmp_path / ("surv_%d.npy" % r)) for r in range(2)])
    assert got.tolist() == alive
    p0, p1 = np.load(tmp_path / "prefix_0.npy"), np.load(tmp_path / "prefix_1.npy")
    assert p0[0] == 0 and p1[0] == len(np.load(tmp_path / "surv_0.npy")) and p0[1] == p1[1] == len(alive)
