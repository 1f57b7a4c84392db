#!/usr/bin/env python3
"""Development aid: device time of every frame of a workload from a cold start (one CUDA-event pair per frame), with the
SM clock and the fraction of (warp, tile) visits that took the checked loop beside it — shows how many frames the world,
the order prediction and the clocks need to reach the steady state the bench reports.

usage: tools/frame_timeline.py [workload] [frames] [flush_mib]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from out_of_nothing_b200 import synth  # noqa: E402
from out_of_nothing_b200.world import World  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c4_cloud_100k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 240
flush = (int(sys.argv[3]) if len(sys.argv) > 3 else 256) << 20

try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    clock = lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)  # noqa: E731
except Exception:
    clock = lambda: -1  # noqa: E731

bodies, cfg = synth.make(workload)
w = World(0)
w.set_props(**cfg["props"])
w.upload(bodies)
chunk = 10
print("# %s, %d bodies, flush %d MiB; columns: first frame of the chunk, mean ms/frame over %d frames, min, max, SM MHz, checked fraction" % (
    workload, bodies.shape[0], flush >> 20, chunk))
for f0 in range(0, frames, chunk):
    ms = [w.bench_frames(cfg["dt"], 1, flush) for _ in range(chunk)]
    ev = w.drain_events()
    print("%4d  %.4f  %.4f  %.4f  %5d  %.3f" % (f0, sum(ms) / chunk, min(ms), max(ms), clock(), ev["warp_tiles_checked"] / max(1, ev["warp_tiles"])), flush=True)
_, prof, launches = w.bench_frames_profiled(cfg["dt"], 20, flush)
print("# steady state per class (ms/frame):", {k: round(v / 20, 4) for k, v in prof.items()}, launches)
w.close()
