"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): one <mu>=10 event through the pixel
path, the strip triplet path, the 100-entry collectors (with and without seedConfirmation, quantised so that the
literal heap replay runs) and the orthogonal seeder.

    compute-sanitizer --tool memcheck python tools/sanitize_driver.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acts_b200 import config, events, plugin  # noqa: E402

mu = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
ev = dict(events.pileup_event(3, mu=mu))
ev["strip"] = events.strip_details(ev, seed=3)
q = dict(ev)
for k in ("x", "y", "z"):
    q[k] = (np.round(ev[k] / np.float32(0.5)) * np.float32(0.5)).astype(np.float32)
q["r"] = np.sqrt(q["x"].astype(np.float64) ** 2 + q["y"].astype(np.float64) ** 2).astype(np.float32)
big = dict(impactWeightFactor=1.0, compatSeedLimit=4, numSeedIncrement=1.0, seedWeightIncrement=10100.0,
           maxSeedsPerSpMConf=100, maxQualitySeedsPerSpMConf=100, maxSeedsPerSpM=4)
cases = [("pixel", {}), ("collectors of 100", big), ("collectors of 100 + confirmation", dict(config.confirmation_overrides(), **big))]
for name, over in cases:
    eng = plugin.SeedingEngine(config.pu200_config(plugin.config_init).update(**over))
    for e in (ev, q):
        a = eng.run(e)
        b = eng.run(e, strip_cot_theta_diff_max=0.2)
        print(name, a["bottom"].size, b["bottom"].size, flush=True)
    eng.close()
ocfg, oopt = config.orthogonal_config(plugin.orthogonal_config_init)
oeng = plugin.SeedingEngine(ocfg, orthogonal=oopt)
print("orthogonal", oeng.run(ev)["bottom"].size)
oeng.close()
