"""Orthogonal (k-d tree) seeder: wall and stage times of a batch of <mu> events on the device, and the unmodified
reference's OrthogonalTripletSeedingAlgorithm on the host cores for comparison.
Usage: python tools/orth_times.py [n_events] [mu] [reps] [--ref]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acts_b200 import config, events, plugin  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_events = int(args[0]) if len(args) > 0 else 8
    mu = float(args[1]) if len(args) > 1 else 200.0
    reps = int(args[2]) if len(args) > 2 else 3
    evs = [events.pileup_event(i, mu=mu) for i in range(n_events)]
    cols, offsets = events.concat_events(evs)
    cfg, opt = config.orthogonal_config(plugin.orthogonal_config_init)
    eng = plugin.SeedingEngine(cfg, orthogonal=opt)
    for rep in range(reps):
        t0 = time.perf_counter()
        res = eng.run_batch(cols, offsets)
        dt = time.perf_counter() - t0
        st = eng.stage_times_ms()
        c = eng.counters()
        print(f"rep {rep}: wall {dt * 1e3:.1f} ms ({n_events / dt:.1f} events/s)  " + "  ".join(f"{k} {v:.2f}" for k, v in st.items()) +
              f"  | launches {c['nKernelLaunches']} seeds {sum(r['bottom'].size for r in res)}", flush=True)
    t0 = time.perf_counter()
    one = eng.run(evs[0])
    print(f"single event: {(time.perf_counter() - t0) * 1e3:.1f} ms, {one['bottom'].size} seeds")
    print({k: v for k, v in eng.counters().items()})
    if "--ref" in sys.argv:
        from oracle import oracle as O
        from oracle import ref as R

        ref = R.Reference(*config.orthogonal_config(O.orthogonal_config_init))
        t0 = time.perf_counter()
        b = ref.run(evs[0])
        dt = time.perf_counter() - t0
        same = all(np.array_equal(one[k].view(np.uint32), b[k].view(np.uint32)) for k in ("bottom", "middle", "top", "quality", "vertexZ"))
        print(f"reference (unmodified sources, one host core): {dt:.2f} s per event, identical seeds: {same}")


if __name__ == "__main__":
    main()
