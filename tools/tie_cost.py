"""Cost of the exact tie-order replays on quantised coordinates (pixel-pitch-like grids: every radius / cotTheta ties).
Events/s of a batch with B200SEED_EXACT_TIES=1 (default, the reference's order) and =0 (canonical order, measurement
only), on smeared events and on events quantised to 50 um / 0.25 mm.  Usage: python tools/tie_cost.py [n_events] [mu]"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def quantise(ev, pitch):
    q = {k: (np.round(v / pitch) * pitch).astype(np.float32) if k in ("x", "y", "z") else v for k, v in ev.items()}
    q["r"] = (np.round(np.hypot(q["x"], q["y"]) / pitch) * pitch).astype(np.float32)
    return q


def child(n_events, mu):
    from acts_b200 import config, events, plugin

    evs = [events.pileup_event(i, mu=mu) for i in range(n_events)]
    for name, pitch in (("smeared", None), ("pitch 0.05 mm", 0.05), ("pitch 0.25 mm", 0.25)):
        batch = evs if pitch is None else [quantise(e, pitch) for e in evs]
        cols, offsets = events.concat_events(batch)
        eng = plugin.SeedingEngine(config.pu200_config(plugin.config_init))
        eng.run_batch(cols, offsets)
        t0 = time.perf_counter()
        for _ in range(3):
            res = eng.run_batch(cols, offsets)
        dt = (time.perf_counter() - t0) / 3
        st, c = eng.stage_times_ms(), eng.counters()
        print(f"EXACT_TIES={os.environ.get('B200SEED_EXACT_TIES', '1')} {name:14s}: {n_events / dt:7.1f} events/s  grid {st['grid']:.2f} ms  "
              f"seed_middles {st['seed_middles']:.1f} ms  tie middles {c['nTieMiddles']} of {c['nMiddles']}  seeds {sum(r['bottom'].size for r in res)}", flush=True)
        eng.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), float(sys.argv[3]))
    else:
        n, mu = (sys.argv[1] if len(sys.argv) > 1 else "8"), (sys.argv[2] if len(sys.argv) > 2 else "200")
        for ties in ("1", "0"):
            env = dict(os.environ, B200SEED_EXACT_TIES=ties)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", n, mu], env=env, check=True)
