"""Replay ACTS CSV space-point dumps through the B200 seeding engine.

    python tools/replay_csv.py --input-dir dump/ --output-dir out/ [--config pu200|seeding_py|itk_like|itk_conf]
                               [--compare]   # compare with event*-seed.csv files found in --input-dir

Reads ``event%09d-spacepoint.csv`` (CsvSpacePointWriter layout; the CsvSpacePointReader layout
``event%09d-spacepoints_pixel.csv`` works too), seeds every event and writes ``event%09d-seed.csv`` in the
CsvSeedWriter layout.  With --compare the (bottom, middle, top) triplets and their quality / vertexZ are matched
against the seed files of the reference run (cross-check against a real ACTS build).
"""
import argparse
import glob
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acts_b200 import config, csvio, plugin  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--input-dir", required=True)
ap.add_argument("--output-dir", required=True)
ap.add_argument("--config", default="pu200", choices=("seeding_py", "pu200", "itk_like", "itk_conf", "itk_pixel", "itk_strip"))
ap.add_argument("--strips", type=float, default=None, metavar="COT_THETA_DIFF_MAX",
                help="strip triplet path (TripletSeedFinder useStripInfo = true) on files that carry the strip columns of the reader layout")
ap.add_argument("--stem", default="spacepoint.csv")
ap.add_argument("--compare", action="store_true")
a = ap.parse_args()

cfg = getattr(config, a.config + "_config")(plugin.config_init)
eng = plugin.SeedingEngine(cfg)
os.makedirs(a.output_dir, exist_ok=True)
files = sorted(glob.glob(os.path.join(a.input_dir, "event*-" + a.stem)))
if not files:
    sys.exit("no event*-%s under %s" % (a.stem, a.input_dir))
bad = 0
for path in files:
    event = int(re.search(r"event(\d+)-", os.path.basename(path)).group(1))
    sp = csvio.read_spacepoints(path)
    if a.strips is not None:
        if "strip" not in sp:
            sys.exit("%s: no strip columns (CsvOutputData.hpp:354-373)" % path)
        seeds = eng.run(sp, strip_cot_theta_diff_max=a.strips)
    else:
        seeds = eng.run(sp)
    params = eng.estimate_params(seeds, sp, b_field=(0.0, 0.0, float(cfg.bFieldInZ)))
    out = csvio.per_event_filepath(a.output_dir, "seed.csv", event)
    csvio.write_seeds(out, seeds, sp, free_params=params, measurement_id=sp["measurement_id"])
    line = "event %d: %d space points -> %d seeds -> %s" % (event, sp["x"].size, seeds["quality"].size, out)
    ref_path = csvio.per_event_filepath(a.input_dir, "seed.csv", event)
    if a.compare and os.path.exists(ref_path):
        ref = csvio.read_seeds(ref_path, measurement_id=sp["measurement_id"])
        mine = {(int(b), int(m), int(t)): (np.float32(q), np.float32(z)) for b, m, t, q, z in
                zip(seeds["bottom"], seeds["middle"], seeds["top"], seeds["quality"], seeds["vertexZ"])}
        theirs = {(int(b), int(m), int(t)): (np.float32(q), np.float32(z)) for b, m, t, q, z in
                  zip(ref["bottom"], ref["middle"], ref["top"], ref["quality"], ref["vertexZ"])}
        common = set(mine) & set(theirs)
        # CsvSeedWriter prints 6 significant digits (default ostream precision)
        differ = sum(1 for k in common if not np.allclose(mine[k], theirs[k], rtol=2e-6, atol=0))
        line += " | reference %d seeds, common %d, only here %d, only there %d, value mismatches %d" % (
            len(theirs), len(common), len(set(mine) - common), len(set(theirs) - common), differ)
        bad += (len(mine) != len(common)) or (len(theirs) != len(common)) or differ != 0
    print(line)
eng.close()
sys.exit(1 if bad else 0)
