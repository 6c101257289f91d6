"""Generates the committed golden vectors in tests/golden/ from the CPU oracle.

The reference holds no golden vector for this path, so these fixtures are produced
by the oracle restatement (oracle/seeding_oracle.cpp).  Every one of them is
reproduced bit for bit by the UNMODIFIED reference sources compiled under
oracle/_ref (tests/test_reference_pin.py::test_golden_fixtures_match_reference),
i.e. they are reference outputs.  They travel to the GPU box, where the CUDA path
is compared with them without running the oracle or the reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from acts_b200 import config, events  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [
    # name, config, generator, event id, mu
    ("seeding_py_muons_ev0", "seeding_py", "muon", 0, 0.0),
    ("seeding_py_muons_ev1", "seeding_py", "muon", 1, 0.0),
    ("pu200_mu10_ev0", "pu200", "pileup", 0, 10.0),
    ("pu200_mu20_ev3", "pu200", "pileup", 3, 20.0),
    ("itk_like_mu10_ev1", "itk_like", "pileup", 1, 10.0),
    ("itk_conf_mu20_ev2", "itk_conf", "pileup", 2, 20.0),  # seedConfirmation = true
    # the verbatim ITk PIXEL configuration (itk.py:302-560) on the ITk-shaped layout
    ("itk_pixel_mu20_ev0", "itk_pixel", "itk", 0, 20.0),
    ("itk_pixel_grid_mu10_ev1", "itk_pixel_grid", "itk", 1, 10.0),
    ("itk_pixel_ho_mu20_ev2", "itk_pixel_ho", "itk", 2, 20.0),
    # the verbatim ITk STRIP configuration (itk.py:458-506: seedConfirmation, collectors of 100) on the strip-shaped
    # layout; these fixtures also carry the strip calibration details and the seeds of the strip triplet path
    # (TripletSeedFinder useStripInfo = true, cotThetaDiffMax = 0.1) under "strip" / "s_*"
    ("itk_strip_mu40_ev0", "itk_strip", "itkstrip", 0, 40.0),
    ("itk_strip_grid_mu20_ev1", "itk_strip_grid", "itkstrip", 1, 20.0),
]
STRIP_COT_THETA_DIFF_MAX = 0.1

from tests.conftest import make_config  # noqa: E402



def main():
    for name, cfg_name, gen, eid, mu in CASES:
        ev = (events.muon_gun_event(eid) if gen == "muon" else
              events.itk_pileup_event(eid, mu=mu) if gen == "itk" else
              events.itk_strip_event(eid, mu=mu) if gen == "itkstrip" else events.pileup_event(eid, mu=mu))
        res = O.Oracle(make_config(cfg_name, O.config_init)).run(ev, want_grid=True)
        extra = {}
        if "strip" in ev:
            sres = O.Oracle(make_config(cfg_name, O.config_init)).run(ev, strip_cot_theta_diff_max=STRIP_COT_THETA_DIFF_MAX)
            extra = dict(strip=ev["strip"], s_cotThetaDiffMax=np.float32(STRIP_COT_THETA_DIFF_MAX),
                         **{"s_" + k: sres[k] for k in ("bottom", "middle", "top", "quality", "vertexZ")})
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), config=cfg_name,
            x=ev["x"], y=ev["y"], z=ev["z"], r=ev["r"], varZ=ev["varZ"], varR=ev["varR"],
            bottom=res["bottom"], middle=res["middle"], top=res["top"], quality=res["quality"],
            vertexZ=res["vertexZ"], grid_copiedFromIndex=res["grid"]["copiedFromIndex"],
            grid_binBegin=res["grid"]["binBegin"], grid_binEnd=res["grid"]["binEnd"],
            counters=np.array([res["counters"][k] for k in ("nInGrid", "nMiddles", "nBottomDoublets", "nTopDoublets", "nCandidates", "nSeeds")], dtype=np.uint64),
            **extra)
        print(name, ev["x"].size, "space points ->", res["quality"].size, "seeds")


if __name__ == "__main__":
    main()
