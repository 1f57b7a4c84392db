"""The expired-body sweep (/root/reference/src/app/OON.cpp:1089-1095) and its sharded form, as executable models.

The reference erases while it iterates and does not re-test the slot it just erased, so inside a maximal run of
consecutive expired bodies only the 1st, 3rd, 5th ... are removed.  The device computes this with scans
(out_of_nothing_b200/csrc/oon_kernels.cu, k_sweep_*), and a sharded world does it rank by rank with ONE small
all-gather: every rank publishes {bodies, length of the expired run its block ends with, whole block is one run} and
derives from the lower ranks' triples how far into a run its first body is.  These tests pin that algorithm on the CPU
(numpy models of the kernels' arithmetic against the literal erase loop); the kernels themselves are checked against the
oracle on the GPU (tests/test_gpu_parity.py, tests/test_gpu_multi.py).
"""
import numpy as np
import pytest

UNLIMITED = np.float32(-1.0)


def literal_sweep(life, player=0):
    """The reference's loop, verbatim: indices (at the time of removal) and the surviving original indices."""
    alive = list(range(len(life)))
    removed_at = []
    i = player + 1
    while i < len(alive):
        l = life[alive[i]]
        if l != UNLIMITED and l <= 0:
            removed_at.append(i)
            del alive[i]            # erase(i); the element sliding into slot i is NOT re-tested
        i += 1
    return removed_at, alive


def scan_sweep(life, first_tested, carry_in=0):
    """k_sweep_flags + max-scan + k_sweep_keep for one block.  Returns keep flags and the block's run triple."""
    n = len(life)
    idx = np.arange(n)
    expired = (idx >= first_tested) & (life != UNLIMITED) & (life <= 0)
    key = np.where(expired, -1, idx)
    last_ok = np.maximum.accumulate(key) if n else key
    pos = idx - last_ok - 1 + np.where(last_ok < 0, carry_in, 0)
    keep = ~expired | ((pos & 1) == 1)
    tail = int(n - 1 - last_ok[-1]) if n else 0
    whole = bool(last_ok[-1] < 0) if n else True
    return keep, (n, tail, whole)


def sharded_sweep(life, bounds, player=0):
    """Rank g owns life[bounds[g]:bounds[g+1]]; the player lives on rank 0."""
    blocks = [life[bounds[g]:bounds[g + 1]] for g in range(len(bounds) - 1)]
    triples = [scan_sweep(b, player + 1 if g == 0 else 0)[1] for g, b in enumerate(blocks)]      # phase 1 + all-gather
    survivors = []
    for g, b in enumerate(blocks):
        carry = 0
        for h in range(g - 1, -1, -1):                                                            # k_sweep_keep's walk down the ranks
            carry += triples[h][1]
            if not triples[h][2]:
                break
        keep, _ = scan_sweep(b, player + 1 if g == 0 else 0, carry)
        survivors += [bounds[g] + int(i) for i in np.nonzero(keep)[0]]
    return survivors


def random_lifetimes(rng, n, p_expired, run_bias):
    life = np.full(n, UNLIMITED, np.float32)
    i = 1
    while i < n:
        if rng.random() < p_expired:
            run = 1 + int(rng.geometric(1.0 - run_bias)) if run_bias > 0 else 1
            life[i:i + run] = rng.choice(np.array([0.0, -0.01, -3.0], np.float32))
            i += run
        else:
            life[i] = rng.choice(np.array([-1.0, 0.5, 7.0], np.float32))
            i += 1
    return life


@pytest.mark.parametrize("seed", range(12))
def test_scan_form_equals_the_literal_erase_loop(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    life = random_lifetimes(rng, n, 0.3, 0.7)
    removed_at, alive = literal_sweep(life)
    keep, _ = scan_sweep(life, 1)
    assert np.nonzero(keep)[0].tolist() == alive
    # index at the time of removal == number of survivors before it (what oon_debug_fetch_removed reports)
    gone = np.nonzero(~keep)[0]
    assert [int(np.count_nonzero(keep[:i])) for i in gone] == removed_at


@pytest.mark.parametrize("seed", range(20))
def test_sharded_form_equals_the_literal_erase_loop(seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(2, 600))
    life = random_lifetimes(rng, n, 0.35, 0.8)
    nranks = int(rng.choice([2, 4, 8]))
    cuts = np.sort(rng.integers(1, n + 1, nranks - 1))          # ragged blocks, some possibly empty; rank 0 owns the player
    bounds = [0, *cuts.tolist(), n]
    _, alive = literal_sweep(life)
    assert sharded_sweep(life, bounds) == alive


def test_runs_across_several_ranks_and_whole_blocks():
    life = np.full(40, UNLIMITED, np.float32)
    life[5:31] = 0                                               # one run of 26 expired bodies ...
    bounds = [0, 8, 12, 12, 20, 33, 40]                          # ... covering two whole blocks and an empty one
    _, alive = literal_sweep(life)
    assert sharded_sweep(life, bounds) == alive
    assert [i for i in range(5, 31) if i not in alive] == list(range(5, 31, 2))   # 1st, 3rd, 5th ... of the run
