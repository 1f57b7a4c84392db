#!/usr/bin/env python3
"""Development aid: where does the time of a short interact launch go?  Per-CTA timeline (oon_debug_trace_interact) of the
launch one GPU of an f-way sharded C4 world runs, summarised: makespan, SM-slot occupancy over time, ramp-up and drain
losses, cost per tile by CTA length.

usage: tools/cta_trace.py [fraction=8] [frames=6]   (env OON_DEBUG_GPC / OON_DEBUG_TAIL_X2 / OON_DEBUG_SPC / OON_SMALL_TILES apply)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from out_of_nothing_b200 import synth  # noqa: E402
from out_of_nothing_b200.world import World  # noqa: E402

fraction = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
bodies, cfg = synth.make("c4_cloud_100k", n=int(os.environ.get("SWEEP_N", "100000")))
dt = cfg["dt"]
w = World(0); w.set_props(**cfg["props"]); w.upload(bodies)
w.step(dt, 150); w.synchronize()
rows = []
for k in range(frames):
    w.update_intrinsic(dt)
    ms_plain = w.debug_time_interact(dt, fraction, k % fraction)
    ms, tr = w.debug_trace_interact(dt, fraction, k % fraction)
    w.update_pairwise(dt); w.update_global(dt)
    sm = (tr[:, 0] & np.uint64(0xffffffff)).astype(np.int64)
    tiles = (tr[:, 0] >> np.uint64(32)).astype(np.int64)
    t0 = tr[:, 1].astype(np.int64); t1 = tr[:, 2].astype(np.int64)
    base = t0.min()
    t0 -= base; t1 -= base
    span = t1.max()
    busy = (t1 - t0).sum()
    slots = 148 * 4
    # occupancy over time in 20 bins
    edges = np.linspace(0, span, 21)
    occ = [np.clip(np.minimum(t1, edges[i + 1]) - np.maximum(t0, edges[i]), 0, None).sum() / ((edges[i + 1] - edges[i]) * slots) for i in range(20)]
    per_tile = {}
    for L in np.unique(tiles):
        sel = tiles == L
        per_tile[int(L)] = (int(sel.sum()), float((t1 - t0)[sel].mean() / 1e3), float((t1 - t0)[sel].mean() / 1e3 / max(L, 1)))
    rows.append((ms_plain, ms, span / 1e6, busy / (span * slots), t0.max() / 1e3))
    if k == frames - 1:
        print("fraction 1/%d: %d CTAs, launch %.4f ms (%.4f traced), makespan of the CTAs %.4f ms, slot occupancy %.3f, last CTA start at %.1f us" % (
            fraction, len(tr), ms_plain, ms, span / 1e6, busy / (span * slots), t0.max() / 1e3))
        print("occupancy by twentieth of the makespan: " + " ".join("%.2f" % o for o in occ))
        print("CTA length (tiles): count, mean duration us, us per tile")
        for L, (c, d, p) in sorted(per_tile.items()):
            print("   %3d tiles: %5d CTAs  %8.2f us  %6.2f us/tile" % (L, c, d, p))
        dur = (t1 - t0) / 1e3
        for L in np.unique(tiles):
            dd = dur[tiles == L]
            print("   %3d tiles: duration p50 %.1f  p90 %.1f  p99 %.1f  max %.1f us" % (L, np.percentile(dd, 50), np.percentile(dd, 90), np.percentile(dd, 99), dd.max()))
        last = np.argsort(t1)[::-1][:12]
        print("last CTAs to finish (tiles, start us, duration us, end us): " + "  ".join("(%d, %.0f, %.0f, %.0f)" % (tiles[i], t0[i] / 1e3, dur[i], t1[i] / 1e3) for i in last))
        first_idle = np.sort(t1)[max(0, len(t1) - slots)] / 1e3 if len(t1) > slots else 0.0
        print("CTAs still running 10 / 20 / 30 us before the end: %d / %d / %d" % tuple(int(((t0 <= span - x * 1e3) & (t1 > span - x * 1e3)).sum()) for x in (10, 20, 30)))
        ends = np.sort(t1)[::-1]
        print("drain: the last 592 CTAs end over %.1f us; time from the first idle slot to the end: %.1f us" % ((ends[0] - ends[min(591, len(ends) - 1)]) / 1e3,
              (span - np.sort(t1)[max(0, len(t1) - slots)]) / 1e3 if len(t1) > slots else 0))
        work_us = float(((t1 - t0)).sum() / 1e3)
        print("sum of CTA durations / slots = %.1f us (the makespan a perfect schedule of these CTAs would have); actual %.1f us" % (work_us / slots, span / 1e3))
print("mean over %d frames: launch %.4f ms" % (frames, float(np.mean([r[0] for r in rows]))))
w.close()
