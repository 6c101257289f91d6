"""Small driver for ncu: seeds N events (<mu>=200 by default) through the C ABI.

    ncu --set full -k regex:k_seed_middles -c 1 python profiles/profile_driver.py --events 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acts_b200 import config, events, plugin  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--events", type=int, default=1)
ap.add_argument("--mu", type=float, default=200.0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--doublets", action="store_true")
ap.add_argument("--conf", action="store_true", help="seedConfirmation = true (ITk pixel filter block)")
a = ap.parse_args()
cfg = config.pu200_config(plugin.config_init)
if a.conf:
    cfg.update(**config.confirmation_overrides())
eng = plugin.SeedingEngine(cfg)
evs = [events.pileup_event(i, mu=a.mu) for i in range(a.events)]
cols, off = events.concat_events(evs)
for _ in range(a.reps):
    res = eng.run_batch(cols, off)
print("seeds", sum(r["quality"].size for r in res), eng.counters(), eng.stage_times_ms())
if "--doublets" in sys.argv:
    d = eng.debug_doublets()
    gb = d["nDoublets"] * 32 / 1e9
    print("materialised doublets: %d middles, %d doublets, %.2f GB written, count+scan+fill %.3f ms -> %.0f GB/s of doublet writes"
          % (d["nMiddles"], d["nDoublets"], gb, d["gpuMilliseconds"], gb / (d["gpuMilliseconds"] * 1e-3)))
