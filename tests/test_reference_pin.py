"""Pins the oracle (the restatement the GPU path is compared with) to the REFERENCE ITSELF.

``oracle/_ref/libseeding_ref.so`` is the unmodified ``GridTripletSeedingAlgorithm`` + ``Core/src/Seeding``
sources of acts-project/acts, compiled where they lie under ``/root/reference`` by ``oracle/Makefile``
(``make ref``) against stand-in headers for the absent third-party pieces (``oracle/ref_shim``).
Every test here runs the same inputs through the reference's own ``execute()`` and through the oracle and
demands bit-identical seeds IN THE SAME ORDER (bottom, middle, top, quality bits, vertexZ bits).
CPU only; skipped when neither the reference tree nor a prebuilt library is present.
"""
import glob
import os

import numpy as np
import pytest

from tests.conftest import make_config

KEYS = ("bottom", "middle", "top", "quality", "vertexZ")


def _same_bits(a, b):
    return all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in KEYS)


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle

    return oracle


@pytest.fixture(scope="module")
def R(built):
    from oracle import ref

    ref.build()
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    return ref


def test_recipe_compiles_only_sources_under_the_reference_tree():
    """The committed recipe names reference translation units by path; nothing of them is in the repo."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mk = open(os.path.join(here, "oracle", "Makefile")).read()
    block = mk[mk.index("REF_SRCS ="):mk.index("REF_INC")]
    srcs = [ln.strip().rstrip("\\").strip() for ln in block.splitlines()[1:] if ln.strip()]
    assert len(srcs) >= 12 and all(s.startswith("$(REF)/") for s in srcs)
    assert any(s.endswith("GridTripletSeedingAlgorithm.cpp") for s in srcs)
    for name in ("DoubletSeedFinder.cpp", "TripletSeedFinder.cpp", "BroadTripletSeedFilter.cpp", "TripletSeeder.cpp",
                 "CandidatesForMiddleSp.cpp", "CylindricalSpacePointGrid.cpp", "SpacePointGridPhiBinning.cpp"):
        assert any(s.endswith("/" + name) for s in srcs), name
        assert not glob.glob(os.path.join(here, "**", name), recursive=True), f"{name} must not be copied into the repo"


@pytest.mark.parametrize("name,kind,mu,ids", [
    ("seeding_py", "muon", 0, (0, 1, 2, 3)),
    ("pu200", "pileup", 5, (0, 1)),
    ("pu200", "pileup", 20, (0, 1, 2)),
    ("pu200", "pileup", 60, (0, 1)),
    ("itk_like", "pileup", 20, (0, 1)),
    ("itk_like", "pileup", 60, (0,)),
    ("itk_conf", "pileup", 20, (0, 2)),
    ("itk_conf", "pileup", 60, (1,)),
    # the verbatim ITk PIXEL configuration (Python/Examples/python/itk.py:302-560) on ITk-shaped events
    ("itk_pixel", "itk", 20, (0, 1)),
    ("itk_pixel", "itk", 60, (2,)),
    ("itk_pixel_grid", "itk", 60, (3,)),
    ("itk_pixel_ho", "itk", 60, (4,)),
])
def test_oracle_matches_reference(O, R, name, kind, mu, ids):
    from acts_b200 import events

    orc = O.Oracle(make_config(name, O.config_init))
    ref = R.Reference(make_config(name, O.config_init))
    for i in ids:
        ev = (events.muon_gun_event(i) if kind == "muon" else
              events.itk_pileup_event(i, mu=mu) if kind == "itk" else events.pileup_event(i, mu=mu))
        a, b = orc.run(ev), ref.run(ev)
        assert b["bottom"].size > 0
        assert _same_bits(a, b), f"{name} event {i}"


def test_oracle_matches_reference_on_random_configurations(O, R):
    """Differential fuzzing of the restatement against the reference: cuts, z binning, custom navigation,
    per-bin middle ranges, variable middle range, ITk cuts, filter knobs, seed confirmation."""
    from acts_b200 import events
    from tests.fuzz import random_config

    rng = np.random.default_rng(20261017)
    done = 0
    for trial in range(120):
        cfg, o = random_config(rng, O.config_init)
        try:
            orc = O.Oracle(cfg)
        except O.OracleError as e:
            with pytest.raises(R.ReferenceError_) as ei:  # the reference constructor refuses it too, same class
                R.Reference(cfg)
            assert ei.value.code == e.code, (o, str(e), str(ei.value))
            continue
        ref = R.Reference(cfg)
        ev = events.pileup_event(300 + trial, mu=float(rng.choice([3, 10, 25, 45])))
        assert _same_bits(orc.run(ev), ref.run(ev)), (trial, o)
        done += 1
    assert done >= 90


def test_vertex_z_windows_match_reference(O, R):
    """Config::inputVertices / vertexZNSigma / vertexZMargin (GridTripletSeedingAlgorithm.hpp:239-243, .cpp:187-206,
    292-297): the reference builds the windows from the vertices; the oracle from the same (z, var z)."""
    from acts_b200 import events

    rng = np.random.default_rng(5)
    for extra, ns, mg in ((0, 3.0, 0.0), (1, 2.0, 1.5), (0, 0.5, 0.25), (1, 5.0, 0.0)):
        cfg = make_config("pu200", O.config_init).update(useVertexZCuts=1, vertexZNSigma=ns, vertexZMargin=mg, useExtraCuts=extra)
        orc, ref = O.Oracle(cfg), R.Reference(cfg)
        for i, nv in ((0, 1), (1, 7), (2, 60), (3, 0)):
            ev = events.pileup_event(40 + i, mu=30)
            vz = rng.normal(0.0, 55.5, nv)
            vv = rng.uniform(0.05, 4.0, nv) ** 2
            a = orc.run(ev, vertices=(vz, vv))
            b = ref.run(ev, vertices=(vz, vv))
            assert _same_bits(a, b), (extra, ns, mg, nv)
            if nv == 0:  # no vertex: every doublet accepted, and the ITk doublet cut is NOT applied (.cpp:292-297)
                plain = O.Oracle(make_config("pu200", O.config_init).update(useExtraCuts=0)).run(ev)
                if not extra:
                    assert _same_bits(a, plain)
            elif nv == 1:  # one narrow window removes most seeds
                assert a["bottom"].size < 0.5 * O.Oracle(make_config("pu200", O.config_init)).run(ev)["bottom"].size


def test_quantised_coordinates_tie_storm_matches_reference(O, R):
    """Pixel-pitch quantised coordinates: radius, cotTheta, curvature and weight ties everywhere, so the
    three unstable std::ranges::sort calls and the heap order decide the result -- the restated libstdc++
    algorithms must agree with the real ones inside the reference build."""
    from acts_b200 import events

    for name in ("pu200", "itk_conf"):
        orc = O.Oracle(make_config(name, O.config_init))
        ref = R.Reference(make_config(name, O.config_init))
        for i, q in ((0, 0.5), (1, 2.0), (2, 8.0)):
            ev = events.pileup_event(60 + i, mu=25)
            x = np.round(ev["x"] / q) * q
            y = np.round(ev["y"] / q) * q
            z = np.round(ev["z"] / (4 * q)) * (4 * q)
            ev = dict(ev, x=x.astype(np.float32), y=y.astype(np.float32), z=z.astype(np.float32),
                      r=np.hypot(x.astype(np.float64), y.astype(np.float64)).astype(np.float32))
            a, b = orc.run(ev), ref.run(ev)
            assert a["counters"]["nCotTieMiddles"] > 0 and a["counters"]["nRTieBins"] > 0
            assert _same_bits(a, b), (name, q)


def test_full_size_event_matches_reference(O, R):
    """BASELINE.json configs[2] size: one <mu>=200 event (~1e5 space points) through the reference's execute()."""
    from acts_b200 import events

    ev = events.pileup_event(0, mu=200)
    a = O.Oracle(make_config("pu200", O.config_init)).run(ev)
    b = R.Reference(make_config("pu200", O.config_init)).run(ev)
    assert b["bottom"].size > 50_000
    assert _same_bits(a, b)


def test_golden_fixtures_match_reference(R, O):
    """tests/golden/*.npz were written by the oracle (tests/golden/make_golden.py); the reference
    reproduces every one of them, so the fixtures are reference outputs."""
    here = os.path.dirname(os.path.abspath(__file__))
    files = sorted(glob.glob(os.path.join(here, "golden", "*.npz")))
    assert len(files) >= 6
    for f in files:
        g = np.load(f, allow_pickle=False)
        name = str(g["config"])
        ev = {k: g[k] for k in ("x", "y", "z", "r", "varZ", "varR")}
        b = R.Reference(make_config(name, O.config_init)).run(ev)
        want = {k: g[k] for k in KEYS}
        assert _same_bits(want, b), os.path.basename(f)
        if "strip" in g.files:  # the reference's own strip triplet path (ref_run_strips) on the same event
            ev["strip"] = g["strip"]
            b = R.Reference(make_config(name, O.config_init)).run_strips(ev, float(g["s_cotThetaDiffMax"]))
            assert _same_bits({k: g["s_" + k] for k in KEYS}, b), os.path.basename(f) + " (strip path)"


def test_reference_entered_from_many_threads(R, O):
    """execute() is const and entered concurrently by the Sequencer's workers (Sequencer.cpp:472-525):
    eight threads through ONE reference algorithm object give the single-threaded seed counts."""
    from acts_b200 import events

    evs = [events.pileup_event(i, mu=20) for i in range(12)]
    cols, offsets = events.concat_events(evs)
    ref = R.Reference(make_config("pu200", O.config_init))
    counts = ref.run_many(cols, offsets, n_threads=8)
    single = [ref.run(ev)["bottom"].size for ev in evs]
    assert counts.tolist() == single


ORTH_VARIANTS = [
    dict(),
    dict(interactionPointCut=1),
    dict(useExtraCuts=1),
    dict(useVariableMiddleSPRange=1, deltaRMiddleMinSPRange=25.0, deltaRMiddleMaxSPRange=40.0),
    dict(deltaPhiMax=0.025, maxSeedsPerSpM=4, sigmaScattering=2.0),
    dict(deltaPhiMax=0.2, zOutermostLayersMin=-650.0, zOutermostLayersMax=800.0, phiMin=-2.5, phiMax=2.0),
    dict(deltaRMinTop=float("nan"), deltaRMaxTop=float("nan"), deltaRMinBottom=6.0, deltaRMaxBottom=150.0),  # the isnan quirk of .cpp:175-198
    dict(deltaRMinTop=10.0, deltaRMaxTop=120.0, deltaRMinBottom=float("nan"), deltaRMaxBottom=float("nan")),
    dict(collisionRegionMin=-80.0, collisionRegionMax=120.0, cotThetaMax=3.0, deltaZMin=-300.0, deltaZMax=250.0),
    dict(useDeltaRinsteadOfTopRadius=1, compatSeedLimit=3, numSeedIncrement=100.0, impactWeightFactor=100.0),
]


@pytest.mark.parametrize("variant", range(len(ORTH_VARIANTS)))
def test_orthogonal_oracle_matches_reference(O, R, variant):
    """The k-d-tree seeder (OrthogonalTripletSeedingAlgorithm.cpp:101-317 + CylindricalSpacePointKDTree.cpp +
    KDTree.hpp): tree construction order, the four search boxes, unsorted doublet search, both z-direction
    groups per middle -- oracle against the reference's own execute(), bit-identical seeds in order."""
    from acts_b200 import config, events

    over = ORTH_VARIANTS[variant]
    orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init, **over))
    ref = R.Reference(*config.orthogonal_config(O.orthogonal_config_init, **over))
    evs = [events.muon_gun_event(variant), events.pileup_event(variant, mu=5), events.pileup_event(40 + variant, mu=30)]
    if variant in (0, 2, 4):
        evs.append(events.itk_pileup_event(variant, mu=10))
    total = 0
    for k, ev in enumerate(evs):
        a, b = orc.run(ev), ref.run(ev)
        total += b["bottom"].size
        assert _same_bits(a, b), f"variant {variant} event {k}"
    assert total > 0


def test_orthogonal_with_seed_confirmation_and_ties(O, R):
    """seedConfirmation through the k-d-tree seeder (one filter state across both groups of a middle and across
    middles, .cpp:234-238) and quantised coordinates (equal keys in the tree's std::sort / std::partition)."""
    from acts_b200 import config, events

    over = dict(config.confirmation_overrides())
    ev = events.pileup_event(5, mu=30)
    a = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init, **over)).run(ev)
    b = R.Reference(*config.orthogonal_config(O.orthogonal_config_init, **over)).run(ev)
    assert b["bottom"].size > 0 and _same_bits(a, b)
    q = {k: (np.round(v * 4) / 4).astype(np.float32) if k in ("x", "y", "z") else v for k, v in ev.items()}
    q["r"] = np.hypot(q["x"], q["y"]).astype(np.float32)
    q["r"] = (np.round(q["r"] * 2) / 2).astype(np.float32)
    a = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init)).run(q)
    b = R.Reference(*config.orthogonal_config(O.orthogonal_config_init)).run(q)
    assert b["bottom"].size > 0 and _same_bits(a, b)


def test_orthogonal_empty_and_tiny_inputs(O, R):
    from acts_b200 import config, events

    orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init))
    ref = R.Reference(*config.orthogonal_config(O.orthogonal_config_init))
    ev = events.pileup_event(0, mu=5)
    for n in (0, 1, 3, 4, 5, 9, 130):
        sub = {k: v[:n] for k, v in ev.items()}
        assert _same_bits(orc.run(sub), ref.run(sub)), n


# ---------------------------------------------------------------------------
# Strip triplet path (TripletSeedFinder.cpp:164-406).  execute() hard-wires useStripInfo = false (.cpp:315), so the
# reference side is ref_run_strips: the reference's own Core objects driven in execute()'s sequence (oracle/ref_driver.cpp).
# ---------------------------------------------------------------------------
def _strip_event(events, kind, i, mu):
    ev = events.muon_gun_event(i) if kind == "muon" else events.pileup_event(i, mu=mu)
    ev = dict(ev)
    ev["strip"] = events.strip_details(ev, seed=i)
    return ev


def test_strip_driver_glue_equals_execute(O, R):
    """ref_run_strips with useStripInfo = 0 takes the pixel path through the same glue: same seeds as execute()."""
    from acts_b200 import events

    for name, kind, mu, i in (("pu200", "pileup", 20, 0), ("seeding_py", "muon", 0, 1), ("pu200", "pileup", 60, 2)):
        ref = R.Reference(make_config(name, O.config_init))
        ev = _strip_event(events, kind, i, mu)
        a, b = ref.run(ev), ref.run_strips(ev, use_strip_info=False)
        assert a["bottom"].size > 0
        assert _same_bits(a, b), f"{name} event {i}"


@pytest.mark.parametrize("name,kind,mu,ids,over", [
    ("seeding_py", "muon", 0, (0, 1), {}),
    ("pu200", "pileup", 5, (0, 1), {}),
    ("pu200", "pileup", 20, (0, 1, 2), {}),
    ("pu200", "pileup", 40, (3,), {}),
    ("pu200", "pileup", 20, (4, 5), dict(seedConfirmation=1)),
    ("pu200", "pileup", 20, (6,), dict(toleranceParam=0.6, interactionPointCut=1)),
    ("pu200", "pileup", 20, (7,), dict(toleranceParam=3.0, useDeltaRinsteadOfTopRadius=1)),
])
def test_strip_oracle_matches_reference(O, R, name, kind, mu, ids, over):
    from acts_b200 import events

    orc = O.Oracle(make_config(name, O.config_init).update(**over))
    ref = R.Reference(make_config(name, O.config_init).update(**over))
    total = 0
    for i in ids:
        ev = _strip_event(events, kind, i, mu)
        for diff in (float("inf"), 0.4, 0.06, 0.0):
            a, b = orc.run(ev, strip_cot_theta_diff_max=diff), ref.run_strips(ev, diff)
            assert _same_bits(a, b), f"{name} event {i} cotThetaDiffMax {diff}"
            total += b["bottom"].size
        # and the strip result is not the pixel result (the calibration moves the points)
        assert not _same_bits(orc.run(ev, strip_cot_theta_diff_max=float("inf")), orc.run(ev)) or ev["x"].size < 50
    assert total > 0


def test_strip_degenerate_details_match_reference(O, R):
    """All-zero details (scale = 0: 0 / 0 in the calibration), parallel strips, huge tolerance, and coordinates
    quantised so that cotTheta values tie."""
    from acts_b200 import events

    orc = O.Oracle(make_config("pu200", O.config_init))
    ref = R.Reference(make_config("pu200", O.config_init))
    ev = _strip_event(events, "pileup", 11, 10)
    zero = dict(ev, strip=np.zeros_like(ev["strip"]))
    par = dict(ev, strip=ev["strip"].copy())
    par["strip"][:, 9:12] = par["strip"][:, 6:9]  # inner strip parallel to the outer one
    q = {k: (np.round(ev[k] * 4) / 4).astype(np.float32) if k in ("x", "y", "z") else ev[k] for k in ev}
    q["r"] = np.hypot(q["x"].astype(np.float64), q["y"].astype(np.float64)).astype(np.float32)
    for case in (zero, par, q):
        for diff in (float("inf"), 0.1):
            a, b = orc.run(case, strip_cot_theta_diff_max=diff), ref.run_strips(case, diff)
            assert _same_bits(a, b)


@pytest.mark.parametrize("conf", [False, True])
def test_itk_strip_filter_block_matches_reference(O, R, conf):
    """Collector capacities 100 / 100 (itk.py:499-506) with and without seedConfirmation, smeared and quantised events."""
    from acts_b200 import config as cm
    from acts_b200 import events

    block = dict(impactWeightFactor=1.0, compatSeedLimit=4, numSeedIncrement=1.0, seedWeightIncrement=10100.0,
                 maxSeedsPerSpMConf=100, maxQualitySeedsPerSpMConf=100, maxSeedsPerSpM=4)
    extra = dict(cm.confirmation_overrides(), **block) if conf else block
    orc = O.Oracle(make_config("pu200", O.config_init).update(**extra))
    ref = R.Reference(make_config("pu200", O.config_init).update(**extra))
    for i, mu, step in ((0, 30, 0.0), (2, 30, 0.5)):
        ev = dict(events.pileup_event(i, mu=mu))
        if step > 0:
            for k in ("x", "y", "z"):
                ev[k] = (np.round(ev[k] / np.float32(step)) * np.float32(step)).astype(np.float32)
            ev["r"] = np.sqrt(ev["x"].astype(np.float64) ** 2 + ev["y"].astype(np.float64) ** 2).astype(np.float32)
        a, b = orc.run(ev), ref.run(ev)
        assert b["bottom"].size > 100
        assert _same_bits(a, b), (conf, i)
