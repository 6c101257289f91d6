"""Stage times of one batch of <mu>=200 events (CUDA events inside the plugin): grid, work list, doublet count,
doublet fill, per-middle seeding, compaction.  Usage: python tools/stage_times.py [n_events] [mu] [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acts_b200 import config, events, plugin  # noqa: E402


def main():
    n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    mu = float(sys.argv[2]) if len(sys.argv) > 2 else 200.0
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    evs = [events.pileup_event(i, mu=mu) for i in range(n_events)]
    cols, offsets = events.concat_events(evs)
    eng = plugin.SeedingEngine(config.pu200_config(plugin.config_init))
    for rep in range(reps):
        t0 = time.perf_counter()
        res = eng.run_batch(cols, offsets)
        dt = time.perf_counter() - t0
        st = eng.stage_times_ms()
        c = eng.counters()
        print(f"rep {rep}: wall {dt * 1e3:.1f} ms  " + "  ".join(f"{k} {v:.2f}" for k, v in st.items()) +
              f"  | launches {c['nKernelLaunches']} seeds {sum(r['bottom'].size for r in res)}", flush=True)
    print({k: v for k, v in eng.counters().items()})


if __name__ == "__main__":
    main()
