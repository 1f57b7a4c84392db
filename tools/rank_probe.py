#!/usr/bin/env python3
"""Development aid (torchrun): per-rank device time per kernel class of a sharded world, with and without the peer-store exchange."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from out_of_nothing_b200 import capi, synth
from out_of_nothing_b200.world import World

rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
bodies, cfg = synth.make(os.environ.get("PROBE_WORKLOAD", "c4_cloud_100k"))
for p2p in (int(x) for x in os.environ.get('PROBE_P2P', '1,0').split(',')):
    os.environ["OON_P2P"] = str(p2p)
    idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    w = World(lr, rank, ws, bytes(idt.cpu().numpy().tobytes()))
    w.set_props(**cfg["props"]); w.upload(bodies)
    w.step(cfg["dt"], 100); w.synchronize()
    dist.barrier()
    w.timer_start(); w.step(cfg["dt"], 50); total = w.timer_stop() / 50
    prof, launches = w.step_profiled(cfg["dt"], 50)
    ftotal, fprof, _ = w.bench_frames_profiled(cfg["dt"], 50, 256 << 20)
    ev = w.drain_events()
    row = {"rank": rank, "kind": w.exchange_kind(), "ms_per_step": round(total, 4), **{k: round(v / 50, 4) for k, v in prof.items()},
           "checked": round(ev["warp_tiles_checked"] / max(1, ev["warp_tiles"]), 3),
           "flushed": {"ms_per_step": round(ftotal / 50, 4), **{k: round(v / 50, 4) for k, v in fprof.items()}}}
    rows = [None] * ws
    dist.all_gather_object(rows, row)
    if rank == 0:
        for r in rows:
            print(json.dumps(r))
    w.close()
dist.destroy_process_group()
