#!/usr/bin/env python
"""Latency mode (BASELINE.json configs[4]): ONE <mu>=300 event split over N GPUs by middle phi sector.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/latency_split.py [--mu 300]

Every rank builds the full grid of the event on its GPU and seeds only its
contiguous block of middle phi bins (b200seed_set_phi_sector); rank 0 gathers
the per-sector seed lists, which concatenate to the unsplit result.  The only
exchange is the host-side gather of the results (no device collective).
Prints one JSON line: per-event latency (max over ranks, CUDA-synchronised wall
clock around b200seed_run) for the split and for one GPU doing the whole event.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from acts_b200 import config, events, plugin, sharding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mu", type=float, default=300.0)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--relaxed", action="store_true", help="the relaxedFloat engine (reported separately)")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = config.pu200_config(plugin.config_init)
cfg.relaxedFloat = 1 if a.relaxed else 0
eng = plugin.SeedingEngine(cfg, device=local)
ev = events.pileup_event(0, mu=a.mu)
# pinned host buffers on both sides, like a production caller (pageable memory costs a staging copy each way)
ev = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in ev.items()}
cap = ev["x"].size * 6
pinned_out = {k: torch.empty(cap, dtype=torch.int32).pin_memory().numpy().view(np.uint32) for k in ("bottom", "middle", "top")}
pinned_out.update({k: torch.empty(cap, dtype=torch.float32).pin_memory().numpy() for k in ("quality", "vertexZ")})


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return out, float(np.median(ts))


first, count = sharding.phi_sector_of_rank(eng.info().phiBins, rank, world)
eng.set_phi_sector(first, count)
mine, t_split = timed(lambda: eng.run(ev, out=pinned_out))
mine = {k: v.copy() for k, v in mine.items()}
eng.set_phi_sector(1, 0)
full, t_full = timed(lambda: eng.run(ev, out=pinned_out))
t = torch.tensor([t_split, t_full], dtype=torch.float64, device="cuda")
parts = [mine]
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v.copy() for k, v in mine.items()})
    parts = gathered
if rank == 0:
    cat = {k: np.concatenate([p[k] for p in parts]) for k in ("bottom", "middle", "top", "quality", "vertexZ")}
    same = all(np.array_equal(cat[k].view(np.uint32), full[k].view(np.uint32)) for k in cat)
    print(json.dumps({"mode": "single-event phi-sector split", "engine": "relaxedFloat" if a.relaxed else "exact", "mu": a.mu, "space_points": int(ev["x"].size),
                      "n_gpus": world, "seeds": int(full["quality"].size), "split_equals_unsplit": bool(same),
                      "latency_ms_split": float(t[0]) * 1e3, "latency_ms_one_gpu": float(t[1]) * 1e3}))
if world > 1:
    dist.destroy_process_group()
