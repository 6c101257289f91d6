#!/usr/bin/env python
"""Benchmark of the seeding hot path: seeded events/s at <mu>=200 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A *step* is one pass of the whole hot path (grid build -> doublets -> triplets
-> filter -> ordered seeds) over one batch of EVENTS_PER_STEP synthetic
<mu>=200 events (~1e5 space points each, Generic-detector-like layout,
acts_b200/events.py).  For N > 1 the script is launched by torchrun, one rank
per GPU; events are independent, so every rank runs its own batches (weak
scaling) and NCCL is only used for the barrier and the max-over-ranks of the
elapsed time.

Printed JSON (one line, rank 0):
  value      events/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        events/s through the host C ABI the way the reference is called: one
             event per b200seed_run call from Sequencer-like worker threads (one
             handle each), pinned host buffers, H2D of the six columns and D2H of
             the seeds inside the timed region (thread count swept to saturation)
  parity     timed events re-seeded by the UNMODIFIED reference (oracle/_ref) on the
             host cores and compared bit for bit, order included; a mismatch fails the run
  roofline   dominant kernel (k_seed_middles, all launches of a step): algorithmic HBM
             bytes / time vs the measured HBM peak; `doublet_stage` is the HBM-bound
             fill pass; `compute` the FP32 view
  cpu_baseline / --impl reference   the reference's own GridTripletSeedingAlgorithm
             (oracle/_ref, built from the unmodified sources) on whole events, one
             event per execute() call from all host threads, like the Sequencer
  latency    config 5: one <mu>=300 event, unsplit on one GPU and split into phi
             sectors over the N ranks of the run
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "seeded events/sec at <mu>=200"
UNIT = "events/s"
WORKLOAD = "generic-detector synthetic pile-up <mu>=200, ~1e5 space points/event, PU200 cut set (sigmaScattering=5)"
N_DISTINCT_EVENTS = 64  # 64 x 2.4 MB of input columns = 154 MB > 126 MB L2
EVENTS_PER_STEP = 16

# algorithmic FP32 operation counts per unit, no FMA
FLOP_PAIR_TEST = 9       # doublet_zr_cuts
FLOP_DOUBLET = 24        # doublet_finish incl. 1 div, 1 sqrt
FLOP_TRIPLET_TEST = 28   # eval_pair, full path incl. 2 div


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.samples.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_events(n, first=0):
    from acts_b200 import events

    return [events.pileup_event(first + i, mu=200.0) for i in range(n)]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------
# reference arm: the reference's own implementation on the host cores
# ---------------------------------------------------------------------------
def reference_engine():
    """(object with run_many(cols, offsets, n_threads), kind): oracle/_ref = the unmodified reference sources
    when the library is there (built where /root/reference is mounted, travels with the snapshot), else the
    oracle port."""
    from acts_b200 import config
    from oracle import oracle as O

    try:
        from oracle import ref as R

        if R.available() or R.build():
            return R.Reference(config.pu200_config(O.config_init)), "reference"
    except Exception:
        pass
    return O.Oracle(config.pu200_config(O.config_init)), "port"


def cpu_reference_rate(n_threads, steps, warmup, evs):
    """events/s of the reference algorithm: one step = n_threads WHOLE events, one event per execute() call,
    handed out to n_threads worker threads that share one algorithm object (Sequencer.cpp:472-525)."""
    from acts_b200 import events

    eng, kind = reference_engine()
    batch = [evs[i % len(evs)] for i in range(n_threads)]
    cols, off = events.concat_events(batch)
    for _ in range(warmup):
        eng.run_many(cols, off, n_threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.run_many(cols, off, n_threads)
    dt = time.perf_counter() - t0
    return steps * n_threads / dt, dt / steps, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    # a whole <mu>=200 event costs ~13 s of one core: one untimed step is warm-up enough for a CPU arm and keeps
    # the run at (1 + K) x ~14 s
    warm = min(args.warmup, 1)
    rate, step_s, kind = cpu_reference_rate(cores, args.steps, warm, make_events(min(cores, 16)))
    sample = (f"{cores} whole events per step, one event per execute() call from {cores} worker threads sharing one "
              f"algorithm object (the Sequencer's pattern); "
              + ("unmodified reference sources (oracle/_ref), default build flags -O2, no -march" if kind == "reference"
                 else "oracle port, -O2, no -march") + f"; {warm} untimed warm-up step(s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_step": cores},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------
KEYS = ("bottom", "middle", "top", "quality", "vertexZ")


def pinned(shape_or_array, dtype=None):
    import torch

    if isinstance(shape_or_array, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(shape_or_array)).pin_memory()
        return t, t.numpy()
    t = torch.empty(shape_or_array, dtype=dtype).pin_memory()
    return t, t.numpy()


def pinned_seed_columns(cap):
    import torch

    keep, out = [], {}
    for k in KEYS:
        t, a = pinned(cap, torch.float32 if k in ("quality", "vertexZ") else torch.int32)
        keep.append(t)
        out[k] = a.view(np.uint32) if k in ("bottom", "middle", "top") else a
    return keep, out


def same_bits(a, b):
    return all(np.array_equal(np.asarray(a[k]).view(np.uint32), np.asarray(b[k]).view(np.uint32)) for k in KEYS)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from acts_b200 import config, events, plugin, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the seeding path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # The contract is ONE JSON line on stdout, but NCCL writes its "NCCL version ..." banner to fd 1 when the first
    # communicator is created: everything between here and the final print goes to stderr instead.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    E = args.events_per_step
    # Weak scaling: every GPU gets the same amount of work per step.  All ranks seed the SAME pool of events (event
    # e of the pool stands for events e, e + pool, ... of a long run): with different random events per rank the
    # max-over-ranks time would measure the spread of the event sizes (4 % at N = 8 in round 1), not the engine.
    n_distinct = max(E, args.distinct_events)
    evs = make_events(n_distinct, first=0)
    cfg = config.pu200_config(plugin.config_init)
    eng = plugin.SeedingEngine(cfg, device=local)
    K = max(1, int(plugin.plan_tables(cfg)["seedsPerMiddle"]))

    # ---- device-resident batches ------------------------------------------
    n_batches = max(1, n_distinct // E)
    batches = []
    for b in range(n_batches):
        cols, off = events.concat_events(evs[b * E:(b + 1) * E])
        n_total = int(off[-1])
        d_cols = [torch.from_numpy(cols[k]).to(dev) for k in ("x", "y", "z", "r", "varZ", "varR")]
        d_off = torch.from_numpy(off.astype(np.int32)).to(dev)
        cap = n_total * K
        d_out = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(3)] + \
                [torch.empty(cap, dtype=torch.float32, device=dev) for _ in range(2)]
        d_soff = torch.zeros(E + 1, dtype=torch.int64, device=dev)
        batches.append(dict(cols=cols, off=off, n_total=n_total, d_cols=d_cols, d_off=d_off, cap=cap, d_out=d_out,
                            d_soff=d_soff))
    # a real (non-NULL) stream: the plugin enqueues on it and the CUDA events below see the work
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step_device(engine, i):
        b = batches[i % n_batches]
        engine.run_batch_device(E, b["n_total"], b["d_off"].data_ptr(), [t.data_ptr() for t in b["d_cols"]],
                                b["d_soff"].data_ptr(), [t.data_ptr() for t in b["d_out"]], b["cap"],
                                stream=C.c_void_p(stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # The timed region: K steps (batches) handed out to `value_streams` worker threads, each with its own handle on its
    # own CUDA stream (the engine synchronises once per call to read its chunk plan; several handles keep the GPU
    # busy across those gaps, like the Sequencer's worker threads do).  Timed with CUDA events: every worker stream
    # waits for the start event, the end event waits for every worker stream.
    n_val = max(1, args.value_streams)
    val_engines = [eng] + [plugin.SeedingEngine(cfg, device=local) for _ in range(n_val - 1)]
    val_streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(n_val - 1)]
    # every worker needs its own output buffers (the batches' inputs are shared, read-only)
    val_out = [[dict(d_out=[torch.empty_like(t) for t in b["d_out"]], d_soff=torch.zeros_like(b["d_soff"])) for b in batches]
               for _ in range(n_val - 1)]

    def step_on(t, i):
        b = batches[i % n_batches]
        o = b if t == 0 else val_out[t - 1][i % n_batches]
        val_engines[t].run_batch_device(E, b["n_total"], b["d_off"].data_ptr(), [x.data_ptr() for x in b["d_cols"]],
                                        o["d_soff"].data_ptr(), [x.data_ptr() for x in o["d_out"]], b["cap"],
                                        stream=C.c_void_p(val_streams[t].cuda_stream))

    def run_steps(first, count):
        nxt = [first]
        lock = threading.Lock()

        def worker(t):
            torch.cuda.set_device(local)
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= first + count:
                    return
                step_on(t, i)

        ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_val)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    run_steps(0, max(args.warmup, n_val) * 1)
    for e_ in val_engines:
        e_.sync()
    launches_per_step = eng.counters()["nKernelLaunches"]

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for st_ in val_streams[1:]:
            st_.wait_event(e0)
        run_steps(args.warmup, args.steps)
        for st_ in val_streams[1:]:
            done = torch.cuda.Event()
            done.record(st_)
            stream.wait_event(done)
        e1.record(stream)
        barrier()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    n_seeds_last = 0
    for e_ in val_engines:
        n_seeds_last = e_.sync() or n_seeds_last
    cnt = eng.counters()
    for e_ in val_engines[1:]:
        e_.close()
    del val_out
    value = world * E * args.steps / (elapsed_ms * 1e-3)
    # per-stage durations, measured live with the plugin's CUDA events (same stream) in a second, identically
    # shaped pass so that reading the events back does not put host syncs into the timed region above
    # (the engine overlaps consecutive arena chunks on two internal streams; a second handle with
    # B200SEED_CHUNK_STREAMS=1 runs them back to back so that every kernel's duration is its own)
    os.environ["B200SEED_CHUNK_STREAMS"] = "1"
    eng_serial = plugin.SeedingEngine(cfg, device=local)
    del os.environ["B200SEED_CHUNK_STREAMS"]
    stage = {k: [] for k in ("grid", "seed", "doublet_count", "doublet_fill", "seed_middles")}
    for i in range(2 + min(args.steps, 8)):
        step_device(eng_serial, args.warmup + i)
        eng_serial.sync()
        if i < 2:
            continue  # workspace allocation
        st = eng_serial.stage_times_ms()
        for k in stage:
            stage[k].append(st[k])
    stage = {k: float(np.mean(v)) for k, v in stage.items()}
    cnt_stage = eng_serial.counters()  # counters of the last batch of the stage pass
    eng_serial.close()

    # ---- end to end through the host C ABI, the reference's call pattern -----------------------------
    # The reference's execute() is entered by several Sequencer worker threads, ONE EVENT PER CALL
    # (Sequencer.cpp:472-525).  T worker threads, each with its own handle and pinned buffers, call the
    # synchronous b200seed_run: the copies of one call overlap the kernels of the others.  T is swept upwards
    # until the rate stops growing.
    ev_pinned = []
    keep_alive = []
    for ev in evs:
        d = {}
        for k in ("x", "y", "z", "r", "varZ", "varR"):
            t, a = pinned(ev[k])
            keep_alive.append(t)
            d[k] = a
        ev_pinned.append(d)
    cap_ev = max(ev["x"].size for ev in evs) * K
    h2d_ev = int(np.mean([ev["x"].size for ev in evs])) * 24

    largest = sorted(range(len(ev_pinned)), key=lambda i: -ev_pinned[i]["x"].size)

    def e2e_per_event(n_thr, n_calls):
        engines = [plugin.SeedingEngine(cfg, device=local) for _ in range(n_thr)]
        outs = [pinned_seed_columns(cap_ev) for _ in range(n_thr)]
        seeds_seen = [0] * n_thr
        nxt = [0]
        lock = threading.Lock()

        def worker(t, limit):
            torch.cuda.set_device(local)
            if limit is None:
                # warm-up: THIS thread's handle seeds the two largest events (the dynamic queue could starve a handle and
                # leave its allocations to the timed region; workspaces grow with the largest event seen)
                for i in largest[:2]:
                    engines[t].run(ev_pinned[i], out=outs[t][1])
                return
            while True:
                with lock:  # dynamic event queue, like tbb::parallel_for over the events
                    i = nxt[0]
                    nxt[0] += 1
                if i >= limit:
                    return
                res = engines[t].run(ev_pinned[i % len(ev_pinned)], out=outs[t][1])
                seeds_seen[t] += res["quality"].size

        def run(limit):
            nxt[0] = 0
            ths = [threading.Thread(target=worker, args=(t, limit)) for t in range(n_thr)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()

        run(None)  # warm-up: every handle allocates its workspaces
        for k in range(n_thr):
            seeds_seen[k] = 0
        barrier()
        t0 = time.perf_counter()
        run(n_calls)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        d2h = 20.0 * sum(seeds_seen) / n_calls
        for e in engines:
            e.close()
        return world * n_calls / dt, d2h

    sweep = {}
    best_t, best_rate, d2h_ev = 1, 0.0, 0.0
    n_calls = max(16, min(64, 4 * args.steps))
    for n_thr in [int(t) for t in args.e2e_threads.split(",")]:
        rate, d2h = e2e_per_event(n_thr, n_calls)
        sweep[str(n_thr)] = rate
        if rate > best_rate:
            best_t, best_rate, d2h_ev = n_thr, rate, d2h
        elif rate < 1.02 * best_rate:
            break  # saturated

    # ---- parity of the timed events against the reference itself ------------------------------------------
    parity = None
    if args.parity_events > 0:
        ref_eng, ref_kind = reference_engine()
        n_check = args.parity_events if world == 1 else min(args.parity_events, 2)
        picks = [(rank * 7 + 5 * j) % n_distinct for j in range(n_check)]  # events of the timed batches
        got_all = {i: eng.run(evs[i]) for i in picks}
        ths = []
        wants = {}

        def ref_worker(i):
            wants[i] = ref_eng.run(evs[i])

        for i in picks:
            th = threading.Thread(target=ref_worker, args=(i,))
            th.start()
            ths.append(th)
        for th in ths:
            th.join()
        bad = sum(0 if same_bits(got_all[i], wants[i]) else 1 for i in picks)
        t = torch.tensor([float(bad), float(len(picks))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        parity = {"events_checked": int(t[1].item()), "mismatches": int(t[0].item()),
                  "against": "unmodified reference sources (oracle/_ref)" if ref_kind == "reference" else "oracle port",
                  "compared": "seed triplets, order, quality bits, vertexZ bits"}

    # ---- config 5: one <mu>=300 event, unsplit and split into phi sectors over the ranks of this run -------
    latency = None
    if args.latency:
        ev300 = events.pileup_event(900, mu=300.0)
        p300 = {}
        for k in ("x", "y", "z", "r", "varZ", "varR"):
            t, a = pinned(ev300[k])
            keep_alive.append(t)
            p300[k] = a
        keep_o, out300 = pinned_seed_columns(ev300["x"].size * K)
        n_phi = eng.info().phiBins

        def timed_run(reps=5):
            eng.run(p300, out=out300)
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                res = eng.run(p300, out=out300)
            dt = (time.perf_counter() - t0) / reps
            return dt, res

        unsplit_s, full = timed_run()
        full = {k: full[k].copy() for k in KEYS}
        latency = {"mu": 300, "space_points": int(ev300["x"].size), "seeds": int(full["quality"].size),
                   "unsplit_ms_one_gpu": unsplit_s * 1e3}
        if world > 1 and world <= n_phi:
            first, count = sharding.phi_sector_of_rank(n_phi, rank, world)
            eng.set_phi_sector(first, count)
            split_s, part = timed_run()
            eng.set_phi_sector(1, 0)
            split_ms = max_over_ranks(split_s) * 1e3
            # the one exchange step: gather the per-sector seed lists (padded) and concatenate in sector order
            n_mine = part["quality"].size
            counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(counts, torch.tensor([n_mine], dtype=torch.int64, device=dev))
            n_max = int(max(int(c.item()) for c in counts))
            gathered = {}
            for k in KEYS:
                mine = torch.zeros(n_max, dtype=torch.int32, device=dev)
                mine[:n_mine] = torch.from_numpy(part[k].view(np.int32).copy()).to(dev)
                parts = [torch.zeros(n_max, dtype=torch.int32, device=dev) for _ in range(world)]
                dist.all_gather(parts, mine)
                gathered[k] = np.concatenate([p_[:int(c.item())].cpu().numpy().view(np.uint32) for p_, c in zip(parts, counts)])
            identical = all(np.array_equal(gathered[k], full[k].view(np.uint32)) for k in KEYS)
            latency.update({"split_ms": split_ms, "sectors": world, "split_equals_unsplit": bool(identical)})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0 if (parity is None or parity["mismatches"] == 0) else 3

    # ---- optional float fast path (relaxedFloat), reported separately -------
    relaxed = None
    if args.relaxed:
        rcfg = config.pu200_config(plugin.config_init)
        rcfg.relaxedFloat = 1
        reng = plugin.SeedingEngine(rcfg, device=local)
        for i in range(args.warmup):
            step_device(reng, i)
        reng.sync()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        r0.record(stream)
        for i in range(args.steps):
            step_device(reng, args.warmup + i)
        r1.record(stream)
        torch.cuda.synchronize()
        reng.sync()
        r_ms = r0.elapsed_time(r1)
        # seed-efficiency delta on one batch: exact engine = the reference's seeds, bit for bit
        b = batches[0]
        ex = eng.run_batch(b["cols"], b["off"], capacity=b["cap"])
        rx = reng.run_batch(b["cols"], b["off"], capacity=b["cap"])
        n_ref = n_rx = n_common = n_same_q = 0
        for a_, b_ in zip(ex, rx):
            sa = {(int(x), int(y), int(z)): float(q) for x, y, z, q in zip(a_["bottom"], a_["middle"], a_["top"], a_["quality"])}
            sb = {(int(x), int(y), int(z)): float(q) for x, y, z, q in zip(b_["bottom"], b_["middle"], b_["top"], b_["quality"])}
            common = set(sa) & set(sb)
            n_ref += len(sa)
            n_rx += len(sb)
            n_common += len(common)
            n_same_q += sum(1 for k in common if sa[k] == sb[k])
        relaxed = {"value": E * args.steps / (r_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "ms_per_step": r_ms / args.steps,
                   "seed_efficiency": n_common / max(1, n_ref), "fake_fraction": (n_rx - n_common) / max(1, n_rx),
                   "seeds_exact": n_ref, "seeds_relaxed": n_rx, "common": n_common, "common_with_identical_quality": n_same_q,
                   "note": "relaxedFloat=1 engine on rank 0 (FMA contraction, approximate division/sqrt, CUDA atan2f, "
                           "no tie replay); efficiency = exact seeds also found / exact seeds on one batch"}
        reng.close()

    # ---- the other TripletSeeder caller: OrthogonalTripletSeedingAlgorithm (k-d-tree provider), reported separately ----
    orthogonal = None
    if args.orthogonal:
        from oracle import ref as R
        from oracle import oracle as O2

        ocfg, oopt = config.orthogonal_config(plugin.orthogonal_config_init)
        oeng = plugin.SeedingEngine(ocfg, device=local, orthogonal=oopt)
        n_o = min(8, E)
        ocols, ooff = events.concat_events(evs[:n_o])
        oeng.run_batch(ocols, ooff)  # warm-up (workspaces)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ores = oeng.run_batch(ocols, ooff)
        o_s = (time.perf_counter() - t0) / reps
        ost = oeng.stage_times_ms()
        ocnt = oeng.counters()
        orth_ref = None
        if R.available() or R.build() is not None:
            rr = R.Reference(*config.orthogonal_config(O2.orthogonal_config_init))
            t0 = time.perf_counter()
            rb = rr.run(evs[0])
            ref_s = time.perf_counter() - t0
            same = all(np.array_equal(ores[0][k].view(np.uint32), rb[k].view(np.uint32)) for k in ("bottom", "middle", "top", "quality", "vertexZ"))
            orth_ref = {"seconds_per_event_one_core": ref_s, "identical_seeds_and_order": bool(same),
                        "kind": "unmodified reference sources (oracle/_ref), one event, one host core"}
            if not same:
                sys.stderr.write("bench.py: PARITY FAILURE of the orthogonal seeder against the reference\n")
        orthogonal = {"value": n_o / o_s, "unit": UNIT, "n_gpus": 1, "events_per_call": n_o, "ms_per_call": o_s * 1e3,
                      "tree_build_ms": ost.get("grid"), "doublet_count_ms": ost.get("doublet_count"),
                      "doublet_fill_ms": ost.get("doublet_fill"), "seed_middles_ms": ost.get("seed_middles"),
                      "seeds": int(sum(r_["bottom"].size for r_ in ores)), "doublets": int(ocnt["nBottomDoublets"] + ocnt["nTopDoublets"]),
                      "reference": orth_ref,
                      "note": "b200seed_create_orthogonal handle, host buffers in / seeds out through b200seed_run_batch "
                              "(wall clock; the k-d trees are built on the device); same <mu>=200 events and cut set"}
        oeng.close()

    # ---- the strip triplet path (TripletSeedFinder useStripInfo = true), reported separately ---------------------
    strips = None
    if args.strips:
        from oracle import ref as R

        cot_diff = 0.05  # cotThetaDiffMax: the pre-filter of TripletSeedFinder.cpp:226-238 (inf = every bottom x top pair)
        seng = plugin.SeedingEngine(config.pu200_config(plugin.config_init), device=local)
        sev = dict(evs[0])
        sev["strip"] = events.strip_details(sev, seed=0)
        seng.run(sev, strip_cot_theta_diff_max=cot_diff)  # warm-up
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            sres = seng.run(sev, strip_cot_theta_diff_max=cot_diff)
        s_s = (time.perf_counter() - t0) / reps
        sst = seng.stage_times_ms()
        scnt = seng.counters()
        strip_ref = None
        if R.available() or R.build() is not None:
            from oracle import oracle as O3

            rr = R.Reference(config.pu200_config(O3.config_init))
            t0 = time.perf_counter()
            rb = rr.run_strips(sev, cot_diff)
            ref_s = time.perf_counter() - t0
            same = all(np.array_equal(sres[k].view(np.uint32), rb[k].view(np.uint32)) for k in ("bottom", "middle", "top", "quality", "vertexZ"))
            strip_ref = {"seconds_per_event_one_core": ref_s, "identical_seeds_and_order": bool(same),
                         "kind": "unmodified reference Core sources driven by oracle/ref_driver.cpp:ref_run_strips, one event, one host core"}
            if not same:
                sys.stderr.write("bench.py: PARITY FAILURE of the strip triplet path against the reference\n")
        strips = {"value": 1.0 / s_s, "unit": UNIT, "n_gpus": 1, "ms_per_event": s_s * 1e3, "cotThetaDiffMax": cot_diff,
                  "seed_middles_ms": sst.get("seed_middles"), "triplet_tests": int(scnt["nTripletTests"]),
                  "seeds": int(sres["bottom"].size), "reference": strip_ref,
                  "note": "b200seed_run_strips, one <mu>=200 event with synthetic double-sided strip module details "
                          "(acts_b200/events.py:strip_details), host buffers in / seeds out (wall clock)"}
        seng.close()

    # ---- roofline ------------------------------------------------------------------------------------------
    peak_gbs, peak_src, sm_max = load_peaks()
    b0 = batches[(args.warmup + min(args.steps, 8) - 1) % n_batches]
    n_in = cnt_stage["nInGrid"]
    n_dbl = cnt_stage["nBottomDoublets"] + cnt_stage["nTopDoublets"]
    # k_seed_middles (dominant: all its launches of one step): every doublet of the arena is read once as a
    # 32-byte record + a 4-byte key, every middle's 32-byte header, 20 bytes per seed slot written
    alg_bytes = 36 * n_dbl + 32 * cnt_stage["nMiddles"] + 20 * cnt_stage["nSeeds"]
    achieved = alg_bytes / (stage["seed_middles"] * 1e-3) / 1e9
    # doublet fill pass (HBM-bound stage): 36 bytes written per doublet + 24 bytes of every packed space point read
    # by its own and the 2 * numPhiNeighbors neighbouring phi bins + the header
    fill_bytes = 36 * n_dbl + (2 * cfg.numPhiNeighbors + 1) * 24 * n_in + 32 * cnt_stage["nMiddles"]
    fill_gbs = fill_bytes / (stage["doublet_fill"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("events_per_step") == E:
            traffic = int(tj["k_seed_middles"]["dram_bytes_read"] + tj["k_seed_middles"]["dram_bytes_write"])
    # FP32 view: operations the reference algorithm itself needs (oracle counts)
    fp32 = None
    if args.oracle_counters:
        from oracle import oracle as O

        orc = O.Oracle(config.pu200_config(O.config_init))
        e_first = ((args.warmup + min(args.steps, 8) - 1) % n_batches) * E
        c = orc.run(evs[e_first])["counters"]
        alg_flop = E * (c["nPairTests"] * FLOP_PAIR_TEST + (c["nBottomDoublets"] + c["nTopDoublets"]) * FLOP_DOUBLET +
                        c["nTripletTests"] * FLOP_TRIPLET_TEST)
        clk0 = clocks.summary()
        sm_mhz = clk0["sm_mhz"] or sm_max
        peak_fp32 = 148 * 128 * sm_mhz * 1e6 / 1e12  # FADD/FMUL per second without FMA, TFLOP/s
        ach = alg_flop / (stage["seed"] * 1e-3) / 1e12
        fp32 = {"bound": "fp32-issue (no FMA allowed on the exact path)", "achieved": ach, "peak": peak_fp32,
                "unit": "TFLOP/s", "frac": ach / peak_fp32,
                "note": "algorithmic flop = oracle pair tests x9 + doublets x24 + triplet tests x28 of the first event of "
                        "the batch x events per step, over the whole seeding stage (doublets + triplets)"}
    clk = clocks.summary()

    cores = host_cores()
    cpu = None
    if not args.no_cpu_baseline:
        rate, step_s, kind = cpu_reference_rate(cores, 1, 0, evs[:min(cores, 16)])
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{cores} whole events, one per execute() call from {cores} worker threads sharing one algorithm "
                         f"object (the Sequencer's pattern), " + ("unmodified reference sources (oracle/_ref)" if kind == "reference" else "oracle port") +
                         f"; {step_s:.1f} s of wall time"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_step_per_gpu": E, "space_points_per_step_per_gpu": b0["n_total"],
                   "l2_policy": f"inputs larger than L2: {n_batches} distinct resident batches "
                                f"({n_batches * b0['n_total'] * 24 / 1e6:.0f} MB of columns) rotated; the doublet arena "
                                f"(~{36 * n_dbl / 1e9:.0f} GB per step) streams through HBM",
                   "parallelism": f"event sharding over {world} GPU(s), same event pool on every rank, no data-path collective",
                   "streams_per_gpu": n_val},
        "clocks": clk,
        "e2e": {"value": best_rate, "unit": UNIT, "h2d_bytes_per_step": int(h2d_ev), "d2h_bytes_per_step": int(d2h_ev),
                "host_threads_per_gpu": best_t, "calls": n_calls, "threads_sweep": sweep,
                "note": "one event per synchronous b200seed_run call (pinned host buffers in, pinned seeds out) from "
                        "host_threads_per_gpu Sequencer-like worker threads, one handle each, dynamic event queue; "
                        "a step of this leg is one event"},
        "gpu_launches": int(launches_per_step * args.steps),
        "parity": parity,
        "roofline": {"bound": "hbm", "kernel": "k_seed_middles", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": stage["seed_middles"],
                     "launch": "all k_seed_middles launches of one step (one per shared-memory class and arena chunk)",
                     "doublet_stage": {"kernel": "k_doublets<fill>", "ms": stage["doublet_fill"], "achieved": fill_gbs,
                                       "frac": fill_gbs / peak_gbs, "algorithmic_bytes": int(fill_bytes),
                                       "count_pass_ms": stage["doublet_count"]},
                     "grid_stage_ms": stage["grid"],
                     "grid_stage_gbs": 52.0 * n_in / (stage["grid"] * 1e-3) / 1e9,
                     "note": "k_seed_middles reads every doublet of the HBM arena once, but it is bound by FP32 "
                             "instruction issue / latency, not by HBM (see `compute`, DESIGN.md section 6); the "
                             "HBM-bound stage is the doublet fill pass (`doublet_stage`)"},
        "compute": fp32,
        "relaxed_float": relaxed,
        "latency": latency,
        "orthogonal": orthogonal,
        "strips": strips,
        "cpu_baseline": cpu,
        "counters_last_step": cnt,
        "seeds_last_step": int(n_seeds_last),
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and parity["mismatches"] != 0:
        sys.stderr.write("bench.py: PARITY FAILURE against the reference on the timed events\n")
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events-per-step", type=int, default=EVENTS_PER_STEP)
    ap.add_argument("--distinct-events", type=int, default=N_DISTINCT_EVENTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--value-streams", type=int, default=3, help="handles (one CUDA stream each) that share the steps of the timed region")
    ap.add_argument("--e2e-threads", default="1,2,3,4,6", help="host worker threads (handles) per GPU swept by the e2e leg")
    ap.add_argument("--parity-events", type=int, default=4, help="timed events re-seeded by the reference and compared (0: skip)")
    ap.add_argument("--no-latency", dest="latency", action="store_false", help="skip the <mu>=300 latency block")
    ap.add_argument("--no-relaxed", dest="relaxed", action="store_false", help="skip the relaxedFloat fast-path report")
    ap.add_argument("--no-oracle-counters", dest="oracle_counters", action="store_false")
    ap.add_argument("--no-orthogonal", dest="orthogonal", action="store_false", help="skip the OrthogonalTripletSeedingAlgorithm report")
    ap.add_argument("--no-strips", dest="strips", action="store_false", help="skip the strip triplet path report")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
