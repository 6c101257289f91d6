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
  e2e        events/s through the host C ABI (b200seed_run_batch) with pinned
             host buffers: H2D of the six columns and D2H of the seeds inside
             the timed region
  roofline   dominant kernel (k_seed_middles): algorithmic HBM bytes / time vs
             the measured HBM peak, plus the FP32 view in `compute` (the fused
             kernel is instruction-issue bound)
  cpu_baseline  the oracle (CPU port of the reference algorithm) timed on the
             host cores on a bounded sample
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
# reference arm: the CPU restatement of the reference algorithm on the host cores
# ---------------------------------------------------------------------------
def cpu_reference_rate(n_threads, steps, warmup, nav_stride, evs=None):
    """events/s of the oracle with n_threads workers.  One step = n_threads
    events, each seeded on 1/nav_stride of its middle bins (bounded sample;
    the grid is built in full), handed out dynamically like the Sequencer's
    parallel_for over events."""
    from acts_b200 import config, events
    from oracle import oracle as O

    orc = O.Oracle(config.pu200_config(O.config_init))
    if evs is None:
        evs = make_events(min(n_threads, 8))
    batch = [evs[i % len(evs)] for i in range(n_threads)]
    cols, off = events.concat_events(batch)
    for _ in range(warmup):
        orc.run_many(cols, off, n_threads, nav_stride)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.run_many(cols, off, n_threads, nav_stride)
    dt = time.perf_counter() - t0
    events_equiv = steps * n_threads / float(nav_stride)
    return events_equiv / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    nav_stride = 8
    rate, step_s = cpu_reference_rate(cores, args.steps, args.warmup, nav_stride)
    sample = (f"{cores} events per step (one per thread), each seeded on every {nav_stride}th middle phi-bin "
              f"(1/{nav_stride} of the event's seeding work, full grid build), oracle C++ -O2 no -march")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_step": cores / nav_stride},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from acts_b200 import config, events, plugin

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
    # every rank seeds its own events (event sharding: event e -> rank e mod world)
    n_distinct = max(E, args.distinct_events)
    evs = make_events(n_distinct, first=rank * n_distinct)
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
        h_cols = {k: torch.from_numpy(cols[k]).pin_memory() for k in cols}
        batches.append(dict(cols=cols, off=off, n_total=n_total, d_cols=d_cols, d_off=d_off, cap=cap, d_out=d_out,
                            d_soff=d_soff, h_cols=h_cols))
    # a real (non-NULL) stream: the plugin enqueues on it and the CUDA events below see the work
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step_device(i):
        b = batches[i % n_batches]
        eng.run_batch_device(E, b["n_total"], b["d_off"].data_ptr(), [t.data_ptr() for t in b["d_cols"]],
                             b["d_soff"].data_ptr(), [t.data_ptr() for t in b["d_out"]], b["cap"],
                             stream=C.c_void_p(stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(i)
    eng.sync()
    launches_per_step = eng.counters()["nKernelLaunches"]

    seed_ms, grid_ms = [], []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for i in range(args.steps):
            step_device(args.warmup + i)
        e1.record(stream)
        barrier()
    elapsed_ms = e0.elapsed_time(e1)
    n_seeds_last = eng.sync()
    cnt = eng.counters()
    # per-kernel durations, measured live with the plugin's CUDA events (same
    # stream) in a second, identically shaped pass so that reading the events
    # back does not put host syncs into the timed region above
    for i in range(min(args.steps, 8)):
        step_device(args.warmup + i)
        eng.sync()
        st = eng.stage_times_ms()
        seed_ms.append(st["seed"])
        grid_ms.append(st["grid"])
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * E * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the host C ABI ---------------------------------
    # The reference runs this algorithm from several Sequencer worker threads, one event (here: batch) per call
    # (Sequencer.cpp:472-525); the plugin's contract is one handle per worker thread.  `e2e_threads` workers, each
    # with its own handle and pinned buffers, call the synchronous b200seed_run_batch: the copies of one call
    # overlap the kernels of the other.
    import threading

    n_thr = max(1, args.e2e_threads)
    engines = [eng] + [plugin.SeedingEngine(cfg, device=local) for _ in range(n_thr - 1)]
    cap_max = max(b["cap"] for b in batches)
    pinned_out = []
    for _ in range(n_thr):
        pinned_out.append({
            "bottom": torch.empty(cap_max, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
            "middle": torch.empty(cap_max, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
            "top": torch.empty(cap_max, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
            "quality": torch.empty(cap_max, dtype=torch.float32).pin_memory().numpy(),
            "vertexZ": torch.empty(cap_max, dtype=torch.float32).pin_memory().numpy()})
    host_cols = [{k: v.numpy() for k, v in b["h_cols"].items()} for b in batches]
    d2h_seen = [0] * n_thr

    def step_host(t, i):
        b = batches[i % n_batches]
        res = engines[t].run_batch(host_cols[i % n_batches], b["off"], out=pinned_out[t])
        d2h_seen[t] = sum(r["quality"].size for r in res) * 20 + (E + 1) * 8

    def worker(t, first, count):
        torch.cuda.set_device(local)
        for i in range(first + t, first + count, n_thr):
            step_host(t, i)

    def run_threads(first, count):
        ths = [threading.Thread(target=worker, args=(t, first, count)) for t in range(n_thr)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    run_threads(0, 2 * n_thr)  # warm-up: every handle allocates its workspaces
    e2e_steps = max(n_thr, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    run_threads(2 * n_thr, e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    d2h = max(d2h_seen)
    for extra in engines[1:]:
        extra.close()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * E * e2e_steps / float(t.item())
    h2d = batches[0]["n_total"] * 24 + (E + 1) * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- optional float fast path (relaxedFloat), reported separately -------
    relaxed = None
    if args.relaxed:
        rcfg = config.pu200_config(plugin.config_init)
        rcfg.relaxedFloat = 1
        reng = plugin.SeedingEngine(rcfg, device=local)

        def step_relaxed(i):
            b = batches[i % n_batches]
            reng.run_batch_device(E, b["n_total"], b["d_off"].data_ptr(), [t.data_ptr() for t in b["d_cols"]],
                                  b["d_soff"].data_ptr(), [t.data_ptr() for t in b["d_out"]], b["cap"],
                                  stream=C.c_void_p(stream.cuda_stream))

        for i in range(args.warmup):
            step_relaxed(i)
        reng.sync()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        r0.record(stream)
        for i in range(args.steps):
            step_relaxed(args.warmup + i)
        r1.record(stream)
        torch.cuda.synchronize()
        reng.sync()
        r_ms = r0.elapsed_time(r1)
        # seed-efficiency delta on one batch: exact engine = the reference's seeds, bit for bit
        b = batches[0]
        host = {k: v.numpy() for k, v in b["h_cols"].items()}
        ex = eng.run_batch(host, b["off"], capacity=b["cap"])
        rx = reng.run_batch(host, b["off"], capacity=b["cap"])
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

    # ---- roofline of the dominant kernel ------------------------------------
    peak_gbs, peak_src, sm_max = load_peaks()
    b0 = batches[(args.warmup + min(args.steps, 8) - 1) % n_batches]
    seed_ms_avg = float(np.mean(seed_ms))
    n_in = cnt["nInGrid"]
    # algorithmic HBM bytes of the FUSED seeding kernel: every
    # packed space point (24 B) is needed by its own and the 2*numPhiNeighbors
    # neighbouring phi bins, plus 8 B per work item and 20 B per seed slot written
    alg_bytes = (2 * cfg.numPhiNeighbors + 1) * 24 * n_in + 8 * cnt["nMiddles"] + 20 * cnt["nSeeds"]
    achieved = alg_bytes / (seed_ms_avg * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed ncu capture (same batch shape)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("events_per_launch") == E:
            traffic = int(tj["dram_bytes_read"] + tj["dram_bytes_write"])
    # FP32 view: operations the reference algorithm itself needs (oracle counts)
    alg_flop = None
    fp32 = None
    if args.oracle_counters:
        from oracle import oracle as O

        orc = O.Oracle(config.pu200_config(O.config_init))
        pair = dbl = trip = 0
        e_first = ((args.warmup + min(args.steps, 8) - 1) % n_batches) * E
        for ev in evs[e_first:e_first + 1]:
            c = orc.run(ev)["counters"]
            pair, dbl, trip = c["nPairTests"], c["nBottomDoublets"] + c["nTopDoublets"], c["nTripletTests"]
        alg_flop = E * (pair * FLOP_PAIR_TEST + dbl * FLOP_DOUBLET + trip * FLOP_TRIPLET_TEST)
    clk = clocks.summary()
    if alg_flop is not None:
        sm_mhz = clk["sm_mhz"] or sm_max
        peak_fp32 = 148 * 128 * sm_mhz * 1e6 / 1e12  # FADD/FMUL per second without FMA, TFLOP/s
        ach = alg_flop / (seed_ms_avg * 1e-3) / 1e12
        fp32 = {"bound": "fp32-issue (no FMA allowed on the exact path)", "achieved": ach, "peak": peak_fp32,
                "unit": "TFLOP/s", "frac": ach / peak_fp32,
                "note": "algorithmic flop = oracle pair tests x9 + doublets x24 + triplet tests x28, first event of the batch x16"}

    cores = host_cores()
    cpu = None
    if not args.no_cpu_baseline:
        nav_stride = 8
        rate, step_s = cpu_reference_rate(cores, 1, 0, nav_stride, evs=evs[:min(cores, 8)])
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cores} events in parallel (one per thread), each seeded on every {nav_stride}th middle "
                         f"phi-bin (1/{nav_stride} of the seeding work, full grid build); {step_s:.1f} s of wall time"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_step_per_gpu": E, "space_points_per_step_per_gpu": b0["n_total"],
                   "l2_policy": f"inputs larger than L2: {n_batches} distinct resident batches "
                                f"({n_batches * b0['n_total'] * 24 / 1e6:.0f} MB of columns) rotated",
                   "parallelism": f"event sharding over {world} GPU(s), no data-path collective"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "host_threads_per_gpu": n_thr, "steps": e2e_steps,
                "note": "synchronous b200seed_run_batch calls (pinned host buffers in, pinned seeds out) from "
                        "host_threads_per_gpu worker threads, one handle each, like the Sequencer's workers"},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": {"bound": "hbm", "kernel": "k_seed_middles", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": seed_ms_avg,
                     "grid_stage_ms": float(np.mean(grid_ms)),
                     "grid_stage_gbs": 52.0 * n_in / (float(np.mean(grid_ms)) * 1e-3) / 1e9,
                     "note": "fused per-middle kernel (all capacity tiers): doublets never leave shared memory, so the "
                             "HBM fraction is small by design and the DRAM traffic (ncu, profiles/r1_traffic.json) is "
                             "below the algorithmic bytes because the packed space points stay in L2; the binding "
                             "limit is FP32 instruction issue and block-barrier latency (see `compute`, DESIGN.md)"},
        "compute": fp32,
        "relaxed_float": relaxed,
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
    ap.add_argument("--e2e-threads", type=int, default=2, help="host worker threads (handles) per GPU of the e2e leg")
    ap.add_argument("--no-relaxed", dest="relaxed", action="store_false", help="skip the relaxedFloat fast-path report")
    ap.add_argument("--no-oracle-counters", dest="oracle_counters", action="store_false")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
