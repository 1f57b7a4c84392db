"""Synthetic world generator: deterministic, rank-independent, the distributions of add_random_body_near (OON.cpp:778-803)."""
import numpy as np

from out_of_nothing_b200 import sharding, synth


def test_counter_based_and_reproducible():
    a, b = synth.random_cloud(5000), synth.random_cloud(5000)
    assert a.tobytes() == b.tobytes()
    big = synth.random_cloud(100000)
    assert big[:5000]["mass"].tobytes() == a["mass"].tobytes()      # body i does not depend on N (positions scale with L only)
    assert synth.random_cloud(5000, seed=1).tobytes() != a.tobytes()


def test_distributions():
    n = 20000
    b = synth.random_cloud(n)
    L = synth.world_extent(n)
    assert b["thrust"][0].tolist() == [0, 0, 0, 0] and (b["thrust"][1:] == np.float32(2e31)).all()
    assert b["mass"][0] == np.float32(1e27) and b["density"][0] == 550
    o = b[1:]
    assert np.abs(o["p"]).max() <= L / 2 and abs(o["p"].mean()) < 0.02 * L
    assert np.abs(o["v"]).max() <= 2.5e8
    assert o["mass"].min() >= np.float32(1e27 / 9) and o["mass"].max() <= np.float32(3e27)
    assert (o["lifetime"] == -1).all() and (o["density"] == 1000).all() and (o["T"] == 0).all()
    assert (o["r"] == synth.radius_from_mass_and_density(o["mass"], o["density"])).all()
    # areal density (bodies per m^2) is the reference's 1000-body value at every N: overlapping partners per body
    # stay O(10), i.e. the colliding-pair FRACTION falls like 1/N (2.6e-4 at 100k bodies)
    R = 2 * float(o["r"].mean())
    assert 10 < n * np.pi * R * R / (L * L) < 40
    assert abs(n / (L * L) - 1000 / (1.5e9 ** 2)) < 1e-20


def test_configs():
    for name in synth.CONFIGS:
        bodies, cfg = synth.make(name, n=3000)
        assert bodies.shape[0] == 3000 and cfg["props"]["interact_n2n"] == 1
    dense, _ = synth.make("c3_dense_13k", n=4000)
    frac = (dense["lifetime"][1:] > 0).mean()
    assert 0.1 < frac < 0.2
    void, _ = synth.make("c5_void_1m", n=4000)
    rad = np.hypot(void["p"][1:, 0], void["p"][1:, 1])
    L = synth.world_extent(4000)
    assert rad.min() >= 0.249 * L and rad.max() <= 0.501 * L
    tang = void["p"][1:, 0] * void["v"][1:, 0] + void["p"][1:, 1] * void["v"][1:, 1]
    assert np.abs(tang / (rad * np.hypot(void["v"][1:, 0], void["v"][1:, 1]))).max() < 1e-3


def test_shards_tile_the_index_space():
    for n in (0, 1, 255, 256, 257, 1010, 10703, 100000, 1000000):
        for g in (1, 2, 3, 4, 8):
            blk = sharding.block_size(n, g)
            assert blk % sharding.TS == 0 and blk * g >= n
            ranges = [sharding.shard_range(n, r, g) for r in range(g)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
    assert sharding.interactions_per_step(1010) == 1010 * 1009
