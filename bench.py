#!/usr/bin/env python3
"""Headless benchmark of the world update (BASELINE.json metric: directed body-pair interactions / s, steps / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4_cloud_100k] [--impl ours|reference]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
one rank per GPU: targets are sharded in contiguous index blocks, the source tiles are exchanged once per frame (peer
stores over NVLink fused into the pack kernel, NCCL all-gather as fallback); torch.distributed only carries the NCCL id,
the barriers and the max-over-ranks.

Workload (default C4 of BASELINE.json, the configuration the metric and the roofline target are quoted on):
100 000-body random cloud, gravity only (friction 0), all-pairs, hyperbolic law, dt 0.033, total size fixed as the
GPU count grows ("strong" scaling).  Rank 0 prints ONE JSON line.

A "step" = one frame: update_intrinsic + update_pairwise + update_global (+ the expired-body sweep when bodies can
expire).  `value` = steps * N*(N-1) / device time of the K timed frames, inputs resident in HBM.  `e2e` = the same
metric through the C ABI with HOST buffers: every frame uploads the body table from pinned host memory, steps once and
reads it back (per-rank blocks with oon_upload_local / oon_download_local on sharded worlds).  `--impl reference` times
the oracle port of the reference's CPU path (the reference itself cannot be built: engine sources absent) on one host
core — the reference is single-threaded by construction.

Beside the headline line (all outside the timed region):
  parity_check  N > 1: KAT-1 x 20 frames in EXACT and in FAST mode, and 3 FAST frames of the 100k-body cloud, on the
                sharded world and on a private 1-GPU world on rank 0's device, compared byte for byte
  c5            N = 8: BASELINE config C5 (1 M bodies, void sphere) on the 8 GPUs + a short 1-GPU leg on rank 0's device
                -> the 8-GPU efficiency of the config the 85 % target is quoted on
  configs       N = 1: C1 (KAT-1), C2, C3 (+ the same world without mortal bodies) in FAST and EXACT mode, C4 in EXACT mode
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_INTERACTION = 18          # literal op count of World.cpp:271-288,304,363-364
INTEGRATE_BYTES_PER_BODY = 77      # BASELINE.md §2 / SURVEY.md §8d
SETTLE_FRAMES = 150                # frames a freshly uploaded world is advanced before the warm-up (see "settle" in the line)
FLUSH_BYTES = 256 << 20            # > 126 MB L2, written on the library's stream before every timed frame


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4_cloud_100k")
    ap.add_argument("--n", type=int, default=0, help="override the body count of the workload")
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--no-l2-flush", action="store_true", help="time the K frames back to back in one enqueue")
    ap.add_argument("--cpu-sample-bodies", type=int, default=0)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--settle", type=int, default=-1, help="frames advanced before the warm-up (default %d for worlds >= 32768 bodies, else 0)" % SETTLE_FRAMES)
    ap.add_argument("--skip-extras", action="store_true", help="headline line only: no parity_check / c5 / configs sub-records")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows, self.proc, self.device_index = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_for_samples(self, n: int, timeout_s: float = 8.0):
        """nvidia-smi needs a second or so to come up on an 8-GPU box: do not start the load before it reports."""
        t0 = time.perf_counter()
        want = len(self.rows) + n
        while self.proc and len(self.rows) < want and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def mark(self):
        return time.perf_counter()

    def stop(self, t_begin: float = 0.0, t_end: float = 1e30) -> dict:
        """Summary of the samples taken under load: from the start of the warm-up (t_begin) to the end of the timed region."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.wait_for_samples(1, 0.5)                      # one more sample right after the timed region
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        rows = [r for t, r in self.rows if t_begin <= t <= t_end + 0.06] or [r for _, r in self.rows[-2:]]
        for row in rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def calibrate_cpu_rate(pyoracle, synth, workload):
    bodies, cfg = synth.make(workload, n=2000)
    o = pyoracle.OracleWorld(double=False)
    o.set_props(**cfg["props"]); o.set_bodies(bodies)
    t = o.time_steps(cfg["dt"], 2)
    return 2 * 2000 * 1999 / t


def run_reference(args):
    """The reference's CPU world update (oracle port: literal restatement, same data layout) on one host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    from out_of_nothing_b200 import synth
    bodies_all, cfg_all = load_workload(args.workload, args.n)
    cfg_n = bodies_all.shape[0]
    rate = calibrate_cpu_rate(pyoracle, synth, args.workload if args.workload in synth.CONFIGS else "c2_swarm_10k")
    budget_s = 90.0
    n_sample = args.cpu_sample_bodies or int(min(cfg_n, max(1000, (budget_s * rate / max(1, args.steps + args.warmup)) ** 0.5)))
    bodies, cfg = bodies_all, cfg_all
    bodies = bodies[:n_sample]
    o = pyoracle.OracleWorld(double=False)
    o.set_props(**cfg["props"]); o.set_bodies(bodies)
    o.step(cfg["dt"], args.warmup)
    n0 = o.n
    secs = o.time_steps(cfg["dt"], args.steps)
    inter = args.steps * n0 * (n0 - 1)
    value = inter / secs
    sample = ("first %d bodies of the %d-body %s world (a full CPU frame of the whole world would take ~%.0f s), %d frames after %d warm-up, "
              "oracle<float> (-O2 -ffp-contract=off): a per-interaction rate on a smaller N, not the same world") % (
        n_sample, cfg_n, args.workload, cfg_n * (cfg_n - 1) / rate, args.steps, args.warmup)
    # the reference's CURRENT number type is double (Cfg.hpp:6; every shipped fixture was written by the float build):
    # a short leg of the double oracle on the same sample, reported beside the float figure
    od = pyoracle.OracleWorld(double=True)
    od.set_props(**cfg["props"]); od.set_bodies(bodies)
    d_steps = 2
    d_secs = od.time_steps(cfg["dt"], d_steps)
    value_f64 = d_steps * n_sample * (n_sample - 1) / d_secs
    line = {
        "impl": "reference", "metric": "body_pair_interactions_per_sec", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "steps_per_sec": args.steps / secs, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "n_bodies": cfg_n, "sample_bodies": n_sample, "dt": cfg["dt"], **cfg["props"]},
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": 1, "kind": "port", "sample": sample,
                         "value_double_build": value_f64, "double_build_sample": "%d frames of oracle<double> on the same bodies" % d_steps},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ our arm
class Ranks:
    """torch.distributed plumbing of the bench: rendezvous, NCCL ids for the library's own communicators, barriers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local_rank)
        if self.size > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def new_world(self, sharded=True):
        """A world handle: one shard of a world over all ranks (its own NCCL communicator), or a private 1-GPU world."""
        from out_of_nothing_b200 import capi
        from out_of_nothing_b200.world import World
        if self.size == 1 or not sharded:
            return World(self.local_rank)
        idt = self.torch.zeros(capi.NCCL_ID_BYTES, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            idt.copy_(self.torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=self.torch.uint8))
        self.dist.broadcast(idt, 0)
        return World(self.local_rank, self.rank, self.size, bytes(idt.cpu().numpy().tobytes()))

    def barrier(self):
        if self.size > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if self.size == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok: bool) -> bool:
        if self.size == 1:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


def load_workload(name, n=None):
    """-> (bodies, cfg).  c1_kat1 = the reference's own tc-smoke start state (tests/golden), the rest is synthetic."""
    from out_of_nothing_b200 import snapshot, synth
    if name == "c1_kat1":
        s = snapshot.load(os.path.join(ROOT, "tests", "golden", "kat1_start.state"))
        props = dict(friction=0.01, gravity_mode=s.gravity_mode, gravity=s.gravity, interact_n2n=1, collision_mode=1,
                     repulsion_stiffness=0.0, loop_mode=0)          # test/regression/tc-smoke.cmd:30-43
        return s.bodies, dict(props=props, dt=0.033, n=s.n, name=name)
    return synth.make(name, n=n or None)


def timed_frames(w, dt, steps, flush_bytes):
    """Device time (ms) of `steps` frames: one event pair per frame with the L2 flushed before each, or one pair around all."""
    if flush_bytes:
        return w.bench_frames(dt, steps, flush_bytes)
    w.timer_start(); w.step(dt, steps)
    return w.timer_stop()


def parity_check(R, capi):
    """Sharded world == private 1-GPU world on rank 0's device, byte for byte (the reference's own byte-compare discipline,
    test/regression/tc-smoke.cmd:30-59).  Both the gathered table (oon_download, every rank) and the shard-local blocks
    (oon_download_local) are compared."""
    import numpy as np
    out = {}
    cases = (("kat1_exact_equals_1gpu", "c1_kat1", None, capi.MATH_EXACT, 20),
             ("kat1_fast_equals_1gpu", "c1_kat1", None, capi.MATH_FAST, 20),
             ("c4_fast_frame_equals_1gpu", "c4_cloud_100k", None, capi.MATH_FAST, 3))
    for key, workload, n, mode, frames in cases:
        bodies, cfg = load_workload(workload, n)
        ws = R.new_world(sharded=True)
        ws.set_props(**cfg["props"]); ws.math_mode = mode
        ws.upload(bodies)
        ws.step(cfg["dt"], frames)
        got = ws.download()                                   # collective: every rank gets the whole table
        first, block = ws.download_local()
        ws.close()
        ok = True
        if R.rank == 0:
            w1 = R.new_world(sharded=False)
            w1.set_props(**cfg["props"]); w1.math_mode = mode
            w1.upload(bodies)
            w1.step(cfg["dt"], frames)
            ref = w1.download()
            w1.close()
            ok = got.tobytes() == ref.tobytes()
            ref_t = R.torch.from_numpy(np.frombuffer(ref.tobytes(), dtype=np.uint8).copy()).cuda()
        else:
            ref_t = R.torch.empty(len(got) * 60, dtype=R.torch.uint8, device="cuda")
        R.dist.broadcast(ref_t, 0)
        ref_all = np.frombuffer(ref_t.cpu().numpy().tobytes(), dtype=capi.BODY_F32)
        ok_local = block.tobytes() == ref_all[first:first + len(block)].tobytes() and got.tobytes() == ref_all.tobytes()
        out[key] = R.all_true(ok and ok_local)
    out["what"] = ("sharded world over %d ranks vs a private 1-GPU world on rank 0's device, same input, same frames; oon_download "
                   "(gathered table, every rank) and oon_download_local (each rank's block) compared byte for byte") % R.size
    return out


def run_c5(R, capi, steps=20, warmup=3):
    """BASELINE config C5 on all ranks + a short 1-GPU leg of the same world on rank 0's device (same lease, same clocks)."""
    bodies, cfg = load_workload("c5_void_1m")
    n, dt = bodies.shape[0], cfg["dt"]
    w = R.new_world(sharded=True)
    w.set_props(**cfg["props"])
    w.upload(bodies)
    w.step(dt, warmup); w.synchronize()
    R.barrier()
    ms = R.max(w.bench_frames(dt, steps, FLUSH_BYTES))
    R.barrier()
    _, prof_ms, prof_l = w.bench_frames_profiled(dt, 5, FLUSH_BYTES)
    k1_ms = R.max(prof_ms["interact"] / max(1, prof_l["interact"]))
    w.close()
    rec = {"workload": "c5_void_1m", "n_bodies": n, "n_gpus": R.size, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
           "value": steps * n * (n - 1) / (ms * 1e-3), "unit": "interactions/s", "interact_ms_per_launch_max_over_ranks": k1_ms}
    one = None
    if R.rank == 0:
        w1 = R.new_world(sharded=False)
        w1.set_props(**cfg["props"])
        w1.upload(bodies)
        w1.step(dt, 1); w1.synchronize()
        ms1 = w1.bench_frames(dt, 3, FLUSH_BYTES) / 3
        peak = capi.measure_fp32_peak(R.local_rank)
        w1.close()
        one = {"ms_per_step": ms1, "value": n * (n - 1) / (ms1 * 1e-3), "steps": 3, "warmup": 1,
               "where": "rank 0's device, the other ranks idle, same process and lease as the %d-GPU leg" % R.size}
        rec["one_gpu"] = one
        rec["efficiency"] = (ms1 / R.size) / (ms / steps)
        rec["interact_frac_of_fp32_peak"] = FLOP_PER_INTERACTION * (n / R.size) * (n - 1) / (k1_ms * 1e-3) * 1e-12 / peak
    R.barrier()
    return rec


def run_configs(R, capi):
    """The other BASELINE configs on one GPU (device time per frame, L2 flushed before every frame): C1 = KAT-1, C2, C3 and the
    same dense world without mortal bodies (what the sweep machinery costs), FAST and EXACT; C4 in EXACT mode."""
    out = {}
    cases = [("c1_kat1", None, False, 200), ("c2_swarm_10k", None, False, 60), ("c3_dense_13k", None, False, 60),
             ("c3_dense_13k", None, True, 60), ("c4_cloud_100k", None, False, 3)]
    for workload, n, immortal, steps in cases:
        bodies, cfg = load_workload(workload, n)
        if immortal:
            bodies = bodies.copy(); bodies["lifetime"] = -1.0
        for mode_name, mode in (("fast", capi.MATH_FAST), ("exact", capi.MATH_EXACT)):
            if workload == "c4_cloud_100k" and mode_name == "fast":
                continue                                     # the headline line itself
            k = steps if mode_name == "fast" or workload == "c1_kat1" else max(3, steps // 6)
            w = R.new_world(sharded=False)
            w.set_props(**cfg["props"]); w.math_mode = mode
            w.upload(bodies)
            w.step(cfg["dt"], 5); w.synchronize()
            n0 = w.entity_count()
            ms = w.bench_frames(cfg["dt"], k, FLUSH_BYTES)
            n1 = w.entity_count()
            ev = w.drain_events()
            w.close()
            nm = 0.5 * (n0 + n1)
            out["%s%s/%s" % (workload, "_immortal" if immortal else "", mode_name)] = {
                "n_bodies": n0, "n_bodies_end": n1, "steps": k, "ms_per_step": ms / k, "steps_per_sec": k / (ms * 1e-3),
                "value": k * nm * (nm - 1) / (ms * 1e-3), "removed": ev["removed"], "terminated": ev["terminated"]}
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np

    from out_of_nothing_b200 import capi, sharding

    R = Ranks(args)
    rank, world_size, local_rank = R.rank, R.size, R.local_rank
    if world_size != args.gpus and world_size == 1 and args.gpus > 1:
        print(json.dumps({"error": "--gpus %d needs torchrun with %d ranks" % (args.gpus, args.gpus)}))
        return 2
    torch = R.torch

    bodies, cfg = load_workload(args.workload, args.n)
    n = bodies.shape[0]
    dt = cfg["dt"]
    w = R.new_world(sharded=True)
    w.set_props(**cfg["props"])
    w.math_mode = capi.MATH_EXACT if args.math == "exact" else capi.MATH_FAST
    w.upload(bodies)

    # ---- live FP32 peak (roofline denominator) before the run, on this device
    fp32_peak = capi.measure_fp32_peak(local_rank)
    flush_bytes = 0 if args.no_l2_flush else FLUSH_BYTES

    # ---- settle + warm-up.  The clock sampler starts here so that short timed regions still get samples under load.
    # The first frames after an upload are not the steady state of the frame loop: the slot order of frame k+1 is predicted
    # from frame k's kicks (none yet), the clocks ramp, and the cloud itself relaxes.  The world is therefore advanced
    # `settle` frames first (reported, timed separately as "first_frames"), then the caller's W warm-up frames, then K timed ones.
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_for_samples(1)
    t_load0 = sampler.mark()
    settle = args.settle if args.settle >= 0 else (SETTLE_FRAMES if (n >= 32768 and args.math == "fast") else 0)
    first_ms = None
    if settle:
        head = min(settle, 5)
        first_ms = R.max(timed_frames(w, dt, head, flush_bytes)) / head
        w.step(dt, settle - head)
    w.step(dt, max(args.warmup, 3))
    w.synchronize()
    n_start = w.entity_count()

    # ---- timed region: device time of exactly K frames, barrier + synchronize on both sides, max over ranks
    R.barrier()
    launches0 = w.launch_count()
    t_host0 = time.perf_counter()
    ms = timed_frames(w, dt, args.steps, flush_bytes)
    R.barrier()
    host_s = time.perf_counter() - t_host0
    clocks = sampler.stop(t_load0, time.perf_counter())
    launches = w.launch_count() - launches0
    ms = R.max(ms)
    n_end = w.entity_count()
    n_mid = (n_start + n_end) // 2                               # bodies of C3 expire on the way: frames in the middle of the region count
    inter_per_step = sharding.interactions_per_step(n_mid)
    value = args.steps * inter_per_step / (ms * 1e-3)

    # ---- per-kernel device time (events on the launching stream) for the roofline lines
    # (same regime as the timed region: L2 flushed before every frame, one event pair per kernel class, no host round trips)
    prof_steps = 20
    if flush_bytes:
        _, prof_ms, prof_launches = w.bench_frames_profiled(dt, prof_steps, flush_bytes)
    else:
        prof_ms, prof_launches = w.step_profiled(dt, prof_steps)
    lo, hi = sharding.shard_range(n_start, rank, world_size)
    local_inter = (hi - lo) * (n_start - 1)
    k1_ms = prof_ms["interact"] / max(1, prof_launches["interact"])
    k1_tflops = FLOP_PER_INTERACTION * local_inter / (k1_ms * 1e-3) * 1e-12 if k1_ms > 0 else 0.0
    k2_ms = prof_ms["integrate"] / max(1, prof_launches["integrate"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    k2_gbs = INTEGRATE_BYTES_PER_BODY * (hi - lo) / (k2_ms * 1e-3) * 1e-9 if k2_ms > 0 else 0.0
    traffic = None
    if world_size == 1 and args.math == "fast":              # one ncu --set full capture of a single-GPU launch; a shard's launch was never captured
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "interact_traffic.json"))).get(args.workload)
        except Exception:
            pass

    # ---- e2e: the drop-in call sequence with HOST buffers (pinned), copies inside the timed region.  Every rank keeps the
    # host table of the reference's `bodies` vector and moves only its own block: oon_upload_local + oon_step + oon_download_local.
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    cur = w.download()
    n_e2e = len(cur)                                         # bodies may have expired since the timed region began (C3)
    mortal_world = bool((cur["lifetime"] != -1.0).any())
    e2e_value = e2e_s = None
    h2d = d2h = 0
    if not (mortal_world and world_size > 1):                # (a sharded world with expiring bodies has ragged blocks: no per-rank host table to cycle)
        host_in = torch.empty(n_e2e * 60, dtype=torch.uint8).pin_memory()
        host_out = torch.empty(n_e2e * 60, dtype=torch.uint8).pin_memory()
        host_in.numpy()[:] = np.frombuffer(cur.tobytes(), dtype=np.uint8)
        first, count = w.local_range(n_e2e)
        state = {"n": n_e2e, "inter": 0}

        def e2e_frame(src, dst):
            n_now = state["n"]
            if world_size > 1:
                w.upload_local_ptr(src.data_ptr() + first * 60, count, n_now)
                w.step(dt, 1)
                w.download_local_ptr(dst.data_ptr() + first * 60, count)
            else:
                w.upload_ptr(src.data_ptr(), n_now)
                w.step(dt, 1)
                state["n"] = w.download_ptr(dst.data_ptr(), n_now)      # the sweep may have removed bodies: the next frame uploads what is left
            state["inter"] += sharding.interactions_per_step(n_now)

        e2e_frame(host_in, host_out)                         # warm (allocations of the staging buffers)
        host_in, host_out = host_out, host_in
        state["inter"] = 0
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_frame(host_in, host_out)
            host_in, host_out = host_out, host_in           # next frame's input is this frame's output
        R.barrier()
        e2e_s = R.max(time.perf_counter() - t0)
        e2e_value = state["inter"] / e2e_s
        h2d = 60 * n_e2e                                     # summed over ranks: every rank uploads its own block ...
        d2h = 60 * n_e2e                                     # ... and reads its own block back

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on one host core, bounded sample
    cpu = None
    if rank == 0 and world_size == 1 and not args.skip_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        ns = args.cpu_sample_bodies or min(n, 30000)
        o = pyoracle.OracleWorld(double=False)
        o.set_props(**cfg["props"]); o.set_bodies(bodies[:ns])
        cpu_steps = 2
        secs = o.time_steps(dt, cpu_steps)
        cpu = {"value": cpu_steps * ns * (ns - 1) / secs, "unit": "interactions/s", "cores": 1, "kind": "port",
               "sample": "first %d bodies of the same world, %d frames, oracle<float> built -O2 -ffp-contract=off, %.1f s" % (ns, cpu_steps, secs)}

    exchange_kind = w.exchange_kind()
    w.close()

    # ---- sub-records outside the timed region
    extras = {}
    if not args.skip_extras and args.workload == "c4_cloud_100k" and args.math == "fast":
        if world_size > 1:
            extras["parity_check"] = parity_check(R, capi)
        if world_size == 8:
            extras["c5"] = run_c5(R, capi)
        if world_size == 1:
            extras["configs"] = run_configs(R, capi)

    if rank == 0:
        line = {
            "metric": "body_pair_interactions_per_sec", "value": value, "unit": "interactions/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "steps_per_sec": args.steps / (ms * 1e-3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_bodies": n, "n_bodies_end": n_end, "dt": dt, "math": args.math,
                       "parallelism": ("targets sharded in %d contiguous blocks; exchange of the source tiles per frame: %s" % (
                           world_size, {"peer_stores": "pack kernel stores each slice into every peer's buffer over NVLink (CUDA IPC), interact kernel waits on per-slice flags",
                                        "nccl_allgather": "NCCL all-gather"}[exchange_kind])) if world_size > 1 else "single GPU",
                       "exchange": exchange_kind,
                       "l2": ("flushed before every timed frame (256 MiB device memset on the same stream, outside the per-frame event pairs)"
                              if flush_bytes else "not flushed; frames timed back to back"),
                       "settle": ("%d frames advanced after the upload and before the %d warm-up frames (the first frames of a fresh world are not the "
                                  "steady state of the frame loop: see first_frames)" % (settle, max(args.warmup, 3))) if settle else "none",
                       **cfg["props"]},
            "first_frames": ({"ms_per_step": first_ms, "value": inter_per_step / (first_ms * 1e-3), "frames": min(settle, 5),
                              "what": "the first frames after the upload, same timing recipe"} if first_ms else None),
            "roofline": {"bound": "fp32", "kernel": "k_interact_fast" if args.math == "fast" else "k_interact_exact",
                         "achieved": k1_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": k1_tflops / fp32_peak if fp32_peak else None, "traffic": traffic,
                         "peak_source": "FFMA microbenchmark measured in this run on this device (MEASURED_PEAKS.json has no FP32 entry)",
                         "algorithmic_flop_per_interaction": FLOP_PER_INTERACTION, "interactions_per_launch": local_inter,
                         "ms_per_launch": k1_ms},
            "roofline_integrate": {"bound": "hbm", "kernel": "k_integrate", "achieved": k2_gbs, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": k2_gbs / hbm_peak, "ms_per_launch": k2_ms,
                                   "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "kernel_ms_per_step": {k: v / prof_steps for k, v in prof_ms.items()},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "interactions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": (e2e_s / e2e_steps * 1e3) if e2e_s else None,
                    "call": ("oon_upload_local(pinned host AoS, own block) + oon_step(dt,1) + oon_download_local(own block) every frame on every rank"
                             if world_size > 1 else "oon_upload(pinned host AoS) + oon_step(dt,1) + oon_download(pinned host AoS) every frame")},
            "gpu_launches": launches,
            "clocks": clocks,
            "host_wall_s_timed_region": host_s,
            **extras,
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        R.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
