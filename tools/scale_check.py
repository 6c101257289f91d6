"""Scale sanity check on a GPU: a 48-event <mu>=200 batch equals the per-event results; seedConfirmation at <mu>=300
and in a 24-event batch (record pool growth, rounds).  python tools/scale_check.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acts_b200 import config, events, plugin
cfg = config.pu200_config(plugin.config_init)
eng = plugin.SeedingEngine(cfg)
evs = [events.pileup_event(i, mu=200) for i in range(48)]
cols, off = events.concat_events(evs)
t = time.time(); res = eng.run_batch(cols, off); print('48-event batch', sum(r['quality'].size for r in res), round(time.time() - t, 2), 's', eng.stage_times_ms())
one = plugin.SeedingEngine(cfg)
for k in (0, 17, 47):
    r = one.run(evs[k])
    assert all(np.array_equal(r[q].view(np.uint32), res[k][q].view(np.uint32)) for q in r), k
print('batch == single ok')
cfgc = config.pu200_config(plugin.config_init).update(**config.confirmation_overrides())
engc = plugin.SeedingEngine(cfgc)
ev = events.pileup_event(3, mu=300)
t = time.time(); r = engc.run(ev); print('mu=300 conf', ev['x'].size, r['quality'].size, engc.counters()['nConfirmationRounds'], round(time.time() - t, 2), 's')
cols, off = events.concat_events(evs[:24])
t = time.time(); resc = engc.run_batch(cols, off); print('24-event conf batch', sum(x['quality'].size for x in resc), round(time.time() - t, 2), 's', engc.counters()['nConfirmationRounds'])
rc = engc.run(evs[5])
assert all(np.array_equal(rc[q].view(np.uint32), resc[5][q].view(np.uint32)) for q in rc)
print('conf batch == single ok')
