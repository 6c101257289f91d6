"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle.

Bit-exact bar: identical (bottom, middle, top) triplets, bit-identical quality
and vertexZ, and -- with the exact tie replay of the bin sort -- the same seed
ORDER as the reference.
"""
import os

import numpy as np
import pytest

from tests.conftest import make_config

pytestmark = pytest.mark.gpu

KEYS = ("bottom", "middle", "top", "quality", "vertexZ")


def _same_bits(a, b):
    return all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in KEYS)


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle

    return oracle


@pytest.fixture(scope="module")
def plugin(built):
    from acts_b200 import plugin

    return plugin


def _event(kind, i, mu):
    from acts_b200 import events

    return events.muon_gun_event(i) if kind == "muon" else events.pileup_event(i, mu=mu)


def test_device_atan2f_matches_host_libm(plugin, O):
    """The device replays glibc's atan2f bit for bit (reference bins with
    std::atan2(float, float), GridTripletSeedingAlgorithm.cpp:219)."""
    rng = np.random.default_rng(7)
    n = 4_000_000
    x = rng.uniform(-200, 200, n).astype(np.float32)
    y = rng.uniform(-200, 200, n).astype(np.float32)
    # special values and raw bit patterns
    x[:64] = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-30] * 8, dtype=np.float32)
    y[:64] = np.repeat(np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-30], dtype=np.float32), 8)
    x[64:100000] = rng.integers(0, 2**32, 100000 - 64, dtype=np.uint32).view(np.float32)
    y[50000:150000] = rng.integers(0, 2**32, 100000, dtype=np.uint32).view(np.float32)
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    dev = eng.device_atan2f(y, x)
    host = np.arctan2(y, x)  # numpy float32 atan2 -> libm atan2f? not guaranteed: use the oracle's libm call on a sample
    sample = rng.choice(n, 200000, replace=False)
    sample[:150] = np.arange(150)
    ref = np.array([O.lib().oracle_atan2f(float(y[i]), float(x[i])) for i in sample], dtype=np.float32)
    got = dev[sample]
    nan = np.isnan(ref) & np.isnan(got)
    assert np.array_equal(got.view(np.uint32)[~nan], ref.view(np.uint32)[~nan])
    del host
    eng.close()


@pytest.mark.parametrize("name,kind,mu,ids", [
    ("seeding_py", "muon", 0, (0, 1, 2)),
    ("pu200", "pileup", 5, (0, 1)),
    ("pu200", "pileup", 20, (0, 1, 2)),
    ("pu200", "pileup", 60, (0, 1)),
    ("itk_like", "pileup", 20, (0, 1)),
    ("itk_like", "pileup", 60, (0,)),
])
def test_seeds_match_oracle(plugin, O, name, kind, mu, ids):
    eng = plugin.SeedingEngine(make_config(name, plugin.config_init))
    orc = O.Oracle(make_config(name, O.config_init))
    for i in ids:
        ev = _event(kind, i, mu)
        got = eng.run(ev)
        ref = orc.run(ev, want_grid=True)
        assert O.seed_set(got) == O.seed_set(ref), f"{name} event {i}: seed set differs"
        assert _same_bits(got, ref), f"{name} event {i}: seed order differs from the reference order"
        cnt = eng.counters()
        assert cnt["nInGrid"] == ref["counters"]["nInGrid"]
        assert cnt["nMiddles"] == ref["counters"]["nMiddles"]
        assert cnt["nBottomDoublets"] == ref["counters"]["nBottomDoublets"]
        assert cnt["nTopDoublets"] == ref["counters"]["nTopDoublets"]
        assert cnt["nCandidates"] == ref["counters"]["nCandidates"]
        # grid stage: packed copy identical incl. the order inside every bin
        g = eng.debug_grid()
        for k in ("copiedFromIndex", "binBegin", "binEnd"):
            assert np.array_equal(g[k], ref["grid"][k]), f"grid {k}"
        for k in ("x", "y", "z", "r", "varZ", "varR"):
            assert np.array_equal(g[k].view(np.uint32), ref["grid"][k].view(np.uint32)), f"grid {k}"
    eng.close()


@pytest.mark.parametrize("words", ["0", "3", "40"])
def test_count_to_fill_hand_off_arena_full_or_off(plugin, O, monkeypatch, words):
    """The count pass hands its r windows and survivor masks to the fill pass through an arena
    (B200SEED_MASK_WORDS_PER_SP words per space point, read at create).  Off (0), nearly always full (3) and
    partly full (40): a middle that did not fit is searched and tested again by the fill pass -- same seeds."""
    monkeypatch.setenv("B200SEED_MASK_WORDS_PER_SP", words)
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    orc = O.Oracle(make_config("pu200", O.config_init))
    for i, mu in ((0, 20), (1, 60)):
        ev = _event("pileup", i, mu)
        got = eng.run(ev)
        ref = orc.run(ev)
        assert _same_bits(got, ref), f"mask words {words}, event {i}"
        cnt = eng.counters()
        assert cnt["nBottomDoublets"] == ref["counters"]["nBottomDoublets"]
        assert cnt["nTopDoublets"] == ref["counters"]["nTopDoublets"]
    eng.close()


@pytest.mark.parametrize("cap", ["520", "800"])
def test_bins_larger_than_the_sort_capacity_use_the_global_scratch(plugin, O, monkeypatch, cap):
    """k_sort_bins orders a bin in shared memory up to B200SEED_SORT_CAP elements (default 2560) and in its slice
    of the global scratch beyond: same packed copy (incl. the order inside bins with equal radii), same seeds."""
    monkeypatch.setenv("B200SEED_SORT_CAP", cap)
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    orc = O.Oracle(make_config("pu200", O.config_init))
    for i, mu, step in ((0, 100, 0.0), (1, 100, 1.0)):
        ev = dict(_event("pileup", i, mu))
        if step > 0:
            for k in ("x", "y", "z"):
                ev[k] = (np.round(ev[k] / np.float32(step)) * np.float32(step)).astype(np.float32)
            ev["r"] = np.sqrt(ev["x"].astype(np.float64) ** 2 + ev["y"].astype(np.float64) ** 2).astype(np.float32)
        got = eng.run(ev)
        ref = orc.run(ev, want_grid=True)
        assert ref["counters"]["maxBinSize"] > int(cap)
        g = eng.debug_grid()
        for k in ("copiedFromIndex", "binBegin", "binEnd"):
            assert np.array_equal(g[k], ref["grid"][k]), f"grid {k}"
        assert _same_bits(got, ref)
    eng.close()


def test_batch_equals_single_events(plugin, O):
    from acts_b200 import events

    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    orc = O.Oracle(make_config("pu200", O.config_init))
    evs = [events.pileup_event(i, mu=m) for i, m in ((0, 20), (1, 5), (2, 40), (3, 1), (4, 20))]
    empty = {k: np.zeros(0, np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")}
    evs.insert(2, empty)
    cols, offsets = events.concat_events(evs)
    res = eng.run_batch(cols, offsets)
    assert len(res) == len(evs)
    for ev, got in zip(evs, res):
        ref = orc.run(ev)
        assert _same_bits(got, ref)
    eng.close()


def _oracle_many(O, cfg_name, evs, overrides=None, **kw):
    """The oracle on several full-size events at once (one thread and one oracle object per event; the
    ctypes call releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    def one(ev):
        cfg = make_config(cfg_name, O.config_init)
        if overrides:
            cfg.update(**overrides)
        return O.Oracle(cfg).run(ev, **kw)

    with ThreadPoolExecutor(max_workers=min(len(evs), os.cpu_count() or 1)) as pool:
        return list(pool.map(one, evs))


def test_eight_full_size_events_mu200(plugin, O):
    """BASELINE.json configs[2] / [3]: eight <mu>=200 events (~1e5 space points each), one call per event and one
    batch call, exact seed sets in the reference's order; the same events with seedConfirmation = true."""
    from acts_b200 import config as cm
    from acts_b200 import events

    evs = [events.pileup_event(10 + i, mu=200) for i in range(8)]
    refs = _oracle_many(O, "pu200", evs)
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    for ev, ref in zip(evs, refs):
        got = eng.run(ev)
        assert ref["quality"].size > 50_000
        assert _same_bits(got, ref)
        cnt = eng.counters()
        for k in ("nMiddles", "nBottomDoublets", "nTopDoublets", "nCandidates"):
            assert cnt[k] == ref["counters"][k], k
    cols, offsets = events.concat_events(evs)
    for got, ref in zip(eng.run_batch(cols, offsets), refs):
        assert _same_bits(got, ref)
    eng.close()
    conf = cm.confirmation_overrides()
    refs = _oracle_many(O, "pu200", evs, overrides=conf)
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init).update(**conf))
    for got, ref in zip(eng.run_batch(cols, offsets), refs):
        assert ref["quality"].size > 10_000
        assert _same_bits(got, ref)
    eng.close()


def test_mu300_event_unsplit_and_eight_sector_split(plugin, O):
    """BASELINE.json configs[4]: one <mu>=300 event (~1.5e5 space points) through the exact engine, unsplit and
    split into 8 phi sectors (run one after the other, as 8 GPUs would in parallel): both bit-identical to the oracle."""
    from acts_b200 import events, sharding

    ev = events.pileup_event(900, mu=300)
    assert ev["x"].size > 130_000
    ref = _oracle_many(O, "pu200", [ev])[0]
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    assert _same_bits(eng.run(ev), ref)
    n_phi = eng.info().phiBins
    parts = []
    for rank in range(8):
        first, count = sharding.phi_sector_of_rank(n_phi, rank, 8)
        eng.set_phi_sector(first, count)
        parts.append(eng.run(ev))
    eng.set_phi_sector(1, 0)
    got = {k: np.concatenate([p[k] for p in parts]) for k in KEYS}
    assert _same_bits(got, ref)
    eng.close()


def test_vertex_z_cuts_match_oracle(plugin, O):
    """Config::inputVertices at the boundary: b200seed_run_vertices builds the windows like .cpp:187-206; the cut is
    connected for every event of such a handle (no vertex: every doublet passes and the ITk doublet cut stays off,
    .cpp:292-297); 200 vertices and per-event windows of a batch are fine (merged intervals, binary search)."""
    from acts_b200 import events

    rng = np.random.default_rng(11)
    for extra, ns, mg in ((0, 3.0, 0.0), (1, 2.0, 1.5)):
        over = dict(useVertexZCuts=1, vertexZNSigma=ns, vertexZMargin=mg, useExtraCuts=extra)
        eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init).update(**over))
        orc = O.Oracle(make_config("pu200", O.config_init).update(**over))
        evs, wins = [], []
        for i, nv in ((0, 1), (1, 7), (2, 200), (3, 0)):
            ev = events.pileup_event(40 + i, mu=30)
            vz, vv = rng.normal(0.0, 55.5, nv), rng.uniform(0.05, 2.0, nv) ** 2
            got = eng.run(ev, vertices=(vz, vv))
            ref = orc.run(ev, vertices=(vz, vv))
            assert _same_bits(got, ref), (extra, nv)
            lo, hi = eng.vertex_windows(vz, vv)
            assert [(float(a), float(b)) for a, b in zip(lo, hi)] == orc.vertex_windows(vz, vv)
            evs.append(ev)
            wins.append(list(zip(lo.tolist(), hi.tolist())))
        cols, offsets = events.concat_events(evs)
        for ev, w, got in zip(evs, wins, eng.run_batch(cols, offsets, z_windows=wins)):
            assert _same_bits(got, orc.run(ev, z_windows=w) if w else orc.run(ev)), extra
        eng.close()


def test_nothing_is_refused_for_its_size(plugin, O):
    """The reference's per-middle containers are unbounded (DoubletSeedFinder.hpp:26-262).  Every phi bin a
    neighbour at <mu>=60 gives middles with ~1e4 doublets per side: far beyond shared memory, taken by the spill
    class (lists in global memory); a tiny arena forces many chunks.  Same seeds, bit for bit."""
    from acts_b200 import events

    ev = events.pileup_event(3, mu=60)
    over = dict(numPhiNeighbors=40)
    ref = O.Oracle(make_config("pu200", O.config_init).update(**over)).run(ev)
    assert ref["counters"]["maxBottoms"] > 5000
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init).update(**over))
    assert _same_bits(eng.run(ev), ref)
    eng.close()
    os.environ["B200SEED_ARENA_MB"] = "64"
    try:
        eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
        ev2 = events.pileup_event(1, mu=200)
        got = eng.run(ev2)
        assert eng.counters()["nKernelLaunches"] > 100  # dozens of arena chunks
        assert _same_bits(got, _oracle_many(O, "pu200", [ev2])[0])
        eng.close()
    finally:
        del os.environ["B200SEED_ARENA_MB"]


def test_full_size_event_mu200(plugin, O):
    """BASELINE.json configs[2]: <mu>=200, ~1e5 space points, exact seed set and order."""
    from acts_b200 import events

    ev = events.pileup_event(0, mu=200)
    assert 80_000 < ev["x"].size < 120_000
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    got = eng.run(ev)
    ref = O.Oracle(make_config("pu200", O.config_init)).run(ev)
    assert O.seed_set(got) == O.seed_set(ref)
    assert _same_bits(got, ref)
    cnt = eng.counters()
    assert cnt["nCandidates"] == ref["counters"]["nCandidates"]
    assert cnt["nTieMiddles"] == ref["counters"]["nCotTieMiddles"]
    eng.close()


def test_seed_confirmation_host_decisions(plugin, O, monkeypatch):
    """The two host-side decisions of the seedConfirmation path: the record pool is grown and the batch re-run when
    the first guess was short, and more rounds are enqueued when the first batch of rounds did not converge."""
    from acts_b200 import config as cm
    from acts_b200 import events

    monkeypatch.setenv("B200SEED_REC_PER_SP", "1")
    monkeypatch.setenv("B200SEED_CONF_ROUNDS", "2")
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init).update(**cm.confirmation_overrides()))
    orc = O.Oracle(make_config("pu200", O.config_init).update(**cm.confirmation_overrides()))
    for i, mu in ((2, 60), (5, 30)):
        ev = events.pileup_event(i, mu=mu)
        got = eng.run(ev)
        ref = orc.run(ev)
        assert ref["counters"]["nCandidates"] > ev["x"].size  # more records than the pool's first guess
        assert _same_bits(got, ref)
        assert eng.counters()["nConfirmationRounds"] > 2      # needed a second batch of rounds
    eng.close()


@pytest.mark.parametrize("name,mu,ids", [
    ("itk_pixel", 20, (0, 1)),
    ("itk_pixel", 60, (2,)),
    ("itk_pixel_grid", 60, (3,)),
    ("itk_pixel_ho", 60, (4,)),
    ("itk_pixel", 200, (5,)),       # ~1.1e5 space points out to |z| = 2.85 m
    ("itk_pixel_ho", 200, (6,)),
])
def test_verbatim_itk_pixel_configuration(plugin, O, name, mu, ids):
    """The reference's second published configuration, verbatim (Python/Examples/python/itk.py:302-560:
    13 z bins, rRangeMiddleSP table, zBinsCustomLooping, interactionPointCut, seedConfirmation, deltaR
    instead of the top radius), on ITk-shaped events; single events and one batch."""
    from acts_b200 import events

    eng = plugin.SeedingEngine(make_config(name, plugin.config_init))
    orc = O.Oracle(make_config(name, O.config_init))
    evs = [events.itk_pileup_event(i, mu=mu) for i in ids]
    refs = [orc.run(ev) for ev in evs]
    for ev, ref in zip(evs, refs):
        got = eng.run(ev)
        assert ref["quality"].size > 0
        assert _same_bits(got, ref), f"{name} mu={mu}"
    if len(evs) > 1:
        cols, offsets = events.concat_events(evs)
        for got, ref in zip(eng.run_batch(cols, offsets), refs):
            assert _same_bits(got, ref)
    eng.close()


def test_full_size_event_mu200_seed_confirmation(plugin, O):
    """<mu>=200 with seedConfirmation = true: ~4.8e4 middles coupled through bestSeedQualityMap."""
    from acts_b200 import config as cm
    from acts_b200 import events

    ev = events.pileup_event(1, mu=200)
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init).update(**cm.confirmation_overrides()))
    got = eng.run(ev)
    ref = O.Oracle(make_config("pu200", O.config_init).update(**cm.confirmation_overrides())).run(ev)
    assert ref["quality"].size > 10_000
    assert _same_bits(got, ref)
    assert eng.counters()["nConfirmationRounds"] >= 2
    eng.close()


@pytest.mark.parametrize("override", [
    dict(interactionPointCut=1),
    dict(useExtraCuts=1),
    dict(useDeltaRinsteadOfTopRadius=1),
    dict(useVariableMiddleSPRange=1),
    dict(maxSeedsPerSpM=0),
    dict(maxSeedsPerSpM=7, maxSeedsPerSpMConf=3),
    dict(compatSeedLimit=0),
    dict(compatSeedLimit=5, seedWeightIncrement=3.5, numSeedIncrement=1.0),
    dict(numPhiNeighbors=2),
    dict(numPhiNeighbors=40),
    dict(phiMin=-1.0, phiMax=2.0),
    dict(deltaZMin=-80.0, deltaZMax=120.0),
    dict(impactMax=10.0, sigmaScattering=50),
    dict(zBinEdges=[-2000.0, -300.0, 0.0, 300.0, 2000.0]),
])
def test_config_variants(plugin, O, override):
    from acts_b200 import events

    cfg = make_config("pu200", plugin.config_init).update(**override)
    cfgo = make_config("pu200", O.config_init).update(**override)
    eng = plugin.SeedingEngine(cfg)
    orc = O.Oracle(cfgo)
    # with every phi bin a neighbour the per-middle lists grow with the whole event: keep it small
    cases = ((0, 3), (5, 6)) if override.get("numPhiNeighbors", 1) > 10 else ((0, 20), (5, 40))
    for i, mu in cases:
        ev = events.pileup_event(i, mu=mu)
        got = eng.run(ev)
        ref = orc.run(ev)
        assert _same_bits(got, ref), f"{override} event {i}"
    eng.close()


@pytest.mark.parametrize("base,extra,cases", [
    ("itk_conf", {}, ((0, 5), (1, 20), (2, 60))),
    ("pu200", "conf", ((0, 20), (3, 60))),
    ("pu200", dict(maxQualitySeedsPerSpMConf=0), ((0, 20),)),
    ("pu200", dict(maxQualitySeedsPerSpMConf=1, maxSeedsPerSpMConf=2, maxSeedsPerSpM=0), ((1, 40),)),
    ("pu200", dict(maxSeedsPerSpM=3, zOriginWeightFactor=0.0, impactWeightFactor=1.0, compatSeedWeight=200.0), ((2, 40),)),
    ("itk_conf", dict(useDeltaRinsteadOfTopRadius=1, seedWeightIncrement=7.0, numSeedIncrement=0.0), ((4, 40),)),
])
def test_seed_confirmation_matches_oracle(plugin, O, base, extra, cases):
    """seedConfirmation = true: the event-wide bestSeedQualityMap order dependence
    (BroadTripletSeedFilter.cpp:278-285,364-376) resolved on the device, bit for bit."""
    from acts_b200 import config as cm
    from acts_b200 import events

    def mk(init):
        cfg = make_config(base, init)
        if base == "pu200":
            cfg.update(**cm.confirmation_overrides())
        if isinstance(extra, dict):
            cfg.update(**extra)
        return cfg

    eng = plugin.SeedingEngine(mk(plugin.config_init))
    orc = O.Oracle(mk(O.config_init))
    for i, mu in cases:
        ev = events.pileup_event(i, mu=mu)
        got = eng.run(ev)
        ref = orc.run(ev)
        assert ref["quality"].size > 0
        assert O.seed_set(got) == O.seed_set(ref), f"{base} {extra} event {i}: seed set differs"
        assert _same_bits(got, ref), f"{base} {extra} event {i}: order differs"
        cnt = eng.counters()
        assert cnt["nCandidates"] == ref["counters"]["nCandidates"]
        assert 2 <= cnt["nConfirmationRounds"] <= 64
    # a batch shares the launches; the fixed point is per event all the same
    evs = [events.pileup_event(10 + k, mu=m) for k, m in enumerate((20, 5, 40))]
    cols, offsets = events.concat_events(evs)
    for ev, got in zip(evs, eng.run_batch(cols, offsets)):
        assert _same_bits(got, orc.run(ev))
    eng.close()


def test_vertex_z_windows(plugin, O):
    """VertexZCuts (GridTripletSeedingAlgorithm.cpp:69-97) takes the experiment-cut slot."""
    from acts_b200 import events

    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init).update(useExtraCuts=1))
    orc = O.Oracle(make_config("pu200", O.config_init).update(useExtraCuts=1))
    ev = events.pileup_event(2, mu=30)
    for windows in ([(-20.0, 15.0)], [(-100.0, -60.0), (-5.0, 5.0), (40.0, 90.0)]):
        got = eng.run(ev, z_windows=windows)
        ref = orc.run(ev, z_windows=windows)
        assert ref["quality"].size > 0
        assert _same_bits(got, ref)
    # no windows again: back to the ITk cuts
    assert _same_bits(eng.run(ev), orc.run(ev))
    eng.close()


def test_edge_cases(plugin, O):
    from acts_b200 import events

    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    orc = O.Oracle(make_config("pu200", O.config_init))
    empty = {k: np.zeros(0, np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")}
    assert eng.run(empty)["quality"].size == 0
    ev = events.pileup_event(9, mu=20)
    # ragged / degenerate inputs: 1, 2, 3 space points
    for n in (1, 2, 3, 17):
        sub = {k: v[:n].copy() for k, v in ev.items()}
        assert _same_bits(eng.run(sub), orc.run(sub))
    # points outside the grid (r >= rMax, |z| >= zMax, NaN) are dropped like the reference does
    bad = {k: v.copy() for k, v in ev.items()}
    bad["r"][::7] = 250.0
    bad["z"][::11] = 2500.0
    bad["z"][::13] = -2000.0  # lower edge is inside
    bad["r"][5] = np.nan
    assert _same_bits(eng.run(bad), orc.run(bad))
    # phi exactly on the closed-axis seam: y = +-0, x < 0 -> atan2f = +-pi
    seam = {k: v.copy() for k, v in ev.items()}
    seam["y"][:50] = 0.0
    seam["y"][50:100] = -0.0
    seam["x"][:100] = -np.abs(seam["x"][:100])
    assert _same_bits(eng.run(seam), orc.run(seam))
    # duplicated space points: equal r and equal cotTheta everywhere -> tie replay
    dup = {k: np.concatenate([v[:3000], v[:3000]]) for k, v in ev.items()}
    got, ref = eng.run(dup), orc.run(dup)
    assert _same_bits(got, ref)
    assert eng.counters()["nTieMiddles"] > 0
    # caller buffer too small -> ERR_CAPACITY with the required size
    from acts_b200 import config as C

    with pytest.raises(plugin.SeedingError) as ei:
        eng.run(ev, capacity=10)
    assert ei.value.code == C.ERR_CAPACITY
    eng.close()


@pytest.mark.parametrize("conf", [False, True])
def test_quantised_coordinates_tie_storm(plugin, O, conf):
    """Coordinates snapped to a coarse lattice: equal radii inside bins, equal cotTheta in the doublet lists, equal
    curvatures in candidate groups and equal weights in the collector heaps -- every order-dependent step of the
    reference (three unstable sorts, the bounded heaps, the strict-greater selections) has to be replayed."""
    from acts_b200 import config as cm
    from acts_b200 import events

    extra = cm.confirmation_overrides() if conf else {}
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init).update(**extra))
    orc = O.Oracle(make_config("pu200", O.config_init).update(**extra))
    for i, mu, step in ((0, 30, 0.5), (1, 60, 2.0)):
        ev = events.pileup_event(i, mu=mu)
        q = {k: v.copy() for k, v in ev.items()}
        for k in ("x", "y", "z"):
            q[k] = (np.round(q[k] / np.float32(step)) * np.float32(step)).astype(np.float32)
        q["r"] = np.sqrt(q["x"].astype(np.float64) ** 2 + q["y"].astype(np.float64) ** 2).astype(np.float32)
        q["varZ"][:] = np.float32(0.01)
        q["varR"][:] = np.float32(0.01)
        got, ref = eng.run(q), orc.run(q)
        assert ref["quality"].size > 100
        assert ref["counters"]["nCotTieMiddles"] > 10 and ref["counters"]["nWeightTieMiddles"] > 0
        assert _same_bits(got, ref), (conf, i)
    eng.close()


def test_canonical_tie_mode_gives_same_seed_set(plugin, O, monkeypatch):
    """B200SEED_EXACT_TIES=0 (canonical (key, index) order): same seed SET on these events."""
    from acts_b200 import events

    monkeypatch.setenv("B200SEED_EXACT_TIES", "0")
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    orc = O.Oracle(make_config("pu200", O.config_init))
    ev = events.pileup_event(1, mu=60)
    assert O.seed_set(eng.run(ev)) == O.seed_set(orc.run(ev))
    eng.close()


def test_gpu_reproduces_committed_golden_vectors(plugin):
    """The CUDA path against the committed fixtures (no oracle involved at run time)."""
    import glob
    import os

    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
    assert len(files) >= 5
    for f in files:
        g = np.load(f)
        ev = {k: g[k] for k in ("x", "y", "z", "r", "varZ", "varR")}
        eng = plugin.SeedingEngine(make_config(str(g["config"]), plugin.config_init))
        got = eng.run(ev)
        for k in KEYS:
            assert np.array_equal(got[k].view(np.uint32), g[k].view(np.uint32)), (f, k)
        grid = eng.debug_grid()
        assert np.array_equal(grid["copiedFromIndex"], g["grid_copiedFromIndex"])
        assert np.array_equal(grid["binBegin"], g["grid_binBegin"]) and np.array_equal(grid["binEnd"], g["grid_binEnd"])
        if "strip" in g.files:  # the strip triplet path on the same event
            ev["strip"] = g["strip"]
            got = eng.run(ev, strip_cot_theta_diff_max=float(g["s_cotThetaDiffMax"]))
            for k in KEYS:
                assert np.array_equal(got[k].view(np.uint32), g["s_" + k].view(np.uint32)), (f, "strip", k)
        eng.close()


def test_property_checks_at_full_size(plugin):
    """Size-independent properties on a full <mu>=200 event (no oracle): per-middle seed cap,
    permutation invariance of the seed SET under a shuffle of the input space points, and
    idempotence (same call twice -> identical output)."""
    from acts_b200 import events

    cfg = make_config("pu200", plugin.config_init)
    eng = plugin.SeedingEngine(cfg)
    ev = events.pileup_event(11, mu=200)
    a = eng.run(ev)
    b = eng.run(ev)
    assert _same_bits(a, b)
    _, counts = np.unique(a["middle"], return_counts=True)
    assert counts.max() <= cfg.maxSeedsPerSpM + 1
    r = ev["r"]
    assert np.all((r[a["middle"]] >= 60.0) & (r[a["middle"]] <= 120.0))
    assert np.all(r[a["bottom"]] < r[a["middle"]]) and np.all(r[a["top"]] > r[a["middle"]])
    rng = np.random.default_rng(3)
    perm = rng.permutation(ev["x"].size)
    shuffled = {k: v[perm] for k, v in ev.items()}
    c = eng.run(shuffled)
    orig = {(int(perm[x]), int(perm[y]), int(perm[z])): (int(q), int(v)) for x, y, z, q, v in
            zip(c["bottom"], c["middle"], c["top"], c["quality"].view(np.uint32), c["vertexZ"].view(np.uint32))}
    ref = {(int(x), int(y), int(z)): (int(q), int(v)) for x, y, z, q, v in
           zip(a["bottom"], a["middle"], a["top"], a["quality"].view(np.uint32), a["vertexZ"].view(np.uint32))}
    # tie order depends on the input order (like in the reference), so allow a handful of differences
    diff = set(orig.items()) ^ set(ref.items())
    assert len(diff) <= 8, len(diff)
    eng.close()


def test_materialised_doublets_match_oracle(plugin, O):
    """Stage-level parity of the PRODUCTION doublet stage (k_doublets, DoubletSeedFinder.cpp:41-273): per middle
    the same bottoms and tops in the reference's emission order with bit-identical
    {cotTheta, iDeltaR, er, u, v, x', y'} -- the 32-byte records the seeding kernel reads from the arena.
    Like the reference (TripletSeeder.cpp:62-69,79) the engine does not search the bottoms of a middle that has
    no tops (or, with seedConfirmation, too few), and drops the tops of a middle without bottoms."""
    from acts_b200 import events

    for name, mu, override in (("pu200", 20, {}), ("itk_like", 20, {}), ("pu200", 40, dict(interactionPointCut=1)),
                               ("itk_conf", 20, {})):
        eng = plugin.SeedingEngine(make_config(name, plugin.config_init).update(**override))
        orc = O.Oracle(make_config(name, O.config_init).update(**override))
        ev = events.pileup_event(2, mu=mu)
        seeds = eng.run(ev)
        got = eng.debug_doublets()
        full = orc.run(ev, dump_doublets=True)
        ref = full["doublets"]
        assert _same_bits(seeds, full)
        assert got["nMiddles"] == ref["middlePos"].size
        assert np.array_equal(got["middlePos"], ref["middlePos"])
        n_checked = 0
        for w in range(got["nMiddles"]):
            g0, g1 = int(got["firstDoublet"][w]), int(got["firstDoublet"][w + 1])
            r0, r1 = int(ref["firstDoublet"][w]), int(ref["firstDoublet"][w + 1])
            if g1 == g0:
                # not searched: the reference's own lists say why (one side empty, or too few tops)
                r_nb = int(ref["nBottom"][w])
                r_nt = (r1 - r0) - r_nb
                assert r_nb == 0 or r_nt == 0 or name == "itk_conf", (name, w, r_nb, r_nt)
                continue
            assert (g1 - g0, int(got["nBottom"][w])) == (r1 - r0, int(ref["nBottom"][w])), (name, w)
            assert np.array_equal(got["otherPos"][g0:g1], ref["otherPos"][r0:r1]), (name, w)
            for k in ("cotTheta", "iDeltaR", "er", "u", "v", "xNew", "yNew"):
                assert np.array_equal(got[k][g0:g1].view(np.uint32), ref[k][r0:r1].view(np.uint32)), (name, w, k)
            n_checked += 1
        assert n_checked > 0.5 * got["nMiddles"]
        cnt = eng.counters()
        assert cnt["nBottomDoublets"] == full["counters"]["nBottomDoublets"]
        assert cnt["nTopDoublets"] == full["counters"]["nTopDoublets"]
        eng.close()


def test_phi_sector_split_concatenates_to_the_full_result(plugin, O):
    """Single-event split (config 5): sectors of middle phi bins, run one after the other (as N GPUs
    would in parallel), concatenate to exactly the unsplit output, for any number of sectors."""
    from acts_b200 import events, sharding

    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    ev = events.pileup_event(3, mu=60)
    ref = O.Oracle(make_config("pu200", O.config_init)).run(ev)
    n_phi = eng.info().phiBins
    for world in (2, 8, 53, 64):  # 64 > 53 phi bins: the surplus ranks hold the empty sector and contribute nothing
        parts = []
        covered = 0
        for rank in range(world):
            first, count = sharding.apply_phi_sector(eng, n_phi, rank, world)
            covered += count
            parts.append(eng.run(ev))
            assert count > 0 or parts[-1]["quality"].size == 0
        assert covered == n_phi
        got = {k: np.concatenate([p[k] for p in parts]) for k in KEYS}
        assert _same_bits(got, ref), world
    eng.set_phi_sector(1, 0)
    assert _same_bits(eng.run(ev), ref)
    eng.close()


def test_estimated_track_parameters_match_reference_arithmetic(plugin, O):
    """'Next' row f1: FP64 free parameters of every seed on the device vs the oracle's restatement of
    Acts::estimateTrackParamsFromSeed.  Tolerance 1e-9 relative (north star), written here."""
    from acts_b200 import events

    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    ev = events.pileup_event(4, mu=60)
    seeds = eng.run(ev)
    assert seeds["quality"].size > 10000
    got = eng.estimate_params(seeds, ev)
    ref = O.estimate_params(seeds, ev)
    assert got.shape == ref.shape == (seeds["quality"].size, 8)
    scale = np.maximum(np.abs(ref), 1e-300)
    rel = np.abs(got - ref) / scale
    # direction components close to zero are compared on the unit-vector scale
    rel[:, 4:7] = np.abs(got[:, 4:7] - ref[:, 4:7])
    assert np.nanmax(rel) < 1e-9, np.nanmax(rel)
    eng.close()


def test_estimate_params_reference_known_answer(plugin, O):
    """Tests/UnitTests/Core/Seeding/EstimateTrackParamsFromSeedTest.cpp:187-195 (trackparm_estimate_aligined):
    three aligned space points give q/p == 0 exactly, on the device and in the oracle."""
    ev = {"x": np.array([-72.775, -84.325, -98.175], np.float32), "y": np.array([-0.325, -0.325, -0.325], np.float32),
          "z": np.array([-615.6, -715.6, -835.6], np.float32)}
    seeds = {"bottom": np.array([0], np.uint32), "middle": np.array([1], np.uint32), "top": np.array([2], np.uint32)}
    b = (0.0, 0.0, 0.000899377)
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    got = eng.estimate_params(seeds, ev, b_field=b)
    ref = O.estimate_params(seeds, ev, b_field=b)
    assert got[0, 7] == 0.0 and ref[0, 7] == 0.0
    assert not np.isnan(got).any()
    # an index outside the space point columns is refused, not read (ADVICE: unchecked device reads)
    bad = dict(seeds, top=np.array([7], np.uint32))
    with pytest.raises(plugin.SeedingError):
        eng.estimate_params(bad, ev, b_field=b)
    eng.close()


def test_csv_replay_round_trip(plugin, O, tmp_path):
    """ACTS CSV dumps -> engine -> CsvSeedWriter layout, and back; equals the oracle on the same file."""
    import subprocess
    import sys

    from acts_b200 import csvio, events

    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    orc = O.Oracle(make_config("pu200", O.config_init))
    refs = {}
    for e, mu in ((0, 5), (7, 20)):
        ev = events.pileup_event(e, mu=mu)
        ids = (np.arange(ev["x"].size, dtype=np.uint64) * 3 + 11)  # measurement ids are not positions
        csvio.write_spacepoints(csvio.per_event_filepath(str(src), "spacepoint.csv", e), ev, measurement_id=ids)
        sp = csvio.read_spacepoints(csvio.per_event_filepath(str(src), "spacepoint.csv", e))
        for k in ("x", "y", "z", "varZ", "varR"):
            assert np.array_equal(sp[k].view(np.uint32), ev[k].view(np.uint32)), k  # 9 digits: exact round trip
        refs[e] = (sp, orc.run(sp))
        # the "reference run" seed file next to the space points, for --compare
        csvio.write_seeds(csvio.per_event_filepath(str(src), "seed.csv", e), refs[e][1], sp, measurement_id=ids)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "replay_csv.py"), "--input-dir", str(src),
                          "--output-dir", str(dst), "--compare"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "only here 0, only there 0, value mismatches 0" in res.stdout
    for e, (sp, ref) in refs.items():
        got = csvio.read_seeds(csvio.per_event_filepath(str(dst), "seed.csv", e), measurement_id=sp["measurement_id"])
        assert np.array_equal(got["bottom"], ref["bottom"]) and np.array_equal(got["top"], ref["top"])
        assert np.array_equal(got["quality"].view(np.uint32), ref["quality"].view(np.uint32))
        assert np.all(got["pT"] > 0)


def test_relaxed_float_fast_path_is_close_but_separate(plugin, O):
    """relaxedFloat = 1 routes to the second engine of the library: almost the same seeds (reported as a
    seed-efficiency delta by bench.py), never claimed exact; the exact engine is untouched by it."""
    from acts_b200 import events

    cfg = make_config("pu200", plugin.config_init)
    cfg.relaxedFloat = 1
    fast = plugin.SeedingEngine(cfg)
    exact = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    orc = O.Oracle(make_config("pu200", O.config_init))
    for i, mu in ((0, 20), (1, 60)):
        ev = events.pileup_event(i, mu=mu)
        ref = orc.run(ev)
        got = fast.run(ev)
        ka, kb = set(O.seed_set(ref)), set(O.seed_set(got))
        eff = len(ka & kb) / len(ka)
        fake = len(kb - ka) / len(kb)
        assert eff > 0.995 and fake < 0.005, (eff, fake)
        assert _same_bits(exact.run(ev), ref)
    assert fast.counters()["nKernelLaunches"] > 0
    fast.close()
    exact.close()
    # the fast path also runs the seedConfirmation rounds; a flipped cut can move a few seeds through the shared map
    from acts_b200 import config as cm

    cfgc = make_config("pu200", plugin.config_init).update(**cm.confirmation_overrides())
    cfgc.relaxedFloat = 1
    fastc = plugin.SeedingEngine(cfgc)
    orcc = O.Oracle(make_config("pu200", O.config_init).update(**cm.confirmation_overrides()))
    ev = events.pileup_event(2, mu=60)
    ka, kb = set(O.seed_set(orcc.run(ev))), set(O.seed_set(fastc.run(ev)))
    assert len(ka & kb) / len(ka) > 0.99 and len(kb - ka) / len(kb) < 0.01
    assert fastc.counters()["nConfirmationRounds"] >= 2
    fastc.close()


def test_pixel_space_points_from_measurements(plugin, O):
    """f4: SpacePointMaker's pixel path on the device (FP64) against the oracle's restatement; the float32 columns
    are what the seeding consumes, so they are compared bit for bit (at most one float ulp where the two double
    evaluation orders round differently) and then seeded."""
    from acts_b200 import events

    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    meas, tr = events.pixel_measurements(1, n=60000)
    got = eng.make_pixel_spacepoints(meas, tr)
    ref = O.make_pixel_spacepoints(meas, tr)
    for k in ("x", "y", "z", "r", "varZ", "varR"):
        assert np.all(np.abs(got[k] - ref[k]) <= np.spacing(np.abs(ref[k]))), k
        assert np.mean(got[k].view(np.uint32) == ref[k].view(np.uint32)) > 0.999, k
    bad = dict(meas, surface=np.full_like(meas["surface"], tr.shape[0]))
    with pytest.raises(plugin.SeedingError):
        eng.make_pixel_spacepoints(bad, tr)
    # the columns feed the seeding call directly
    orc = O.Oracle(make_config("pu200", O.config_init))
    assert _same_bits(eng.run(got), orc.run(got))
    # ... and the fused entry (no host round trip of the space points) gives the same columns and seeds
    seeds, sp = eng.run_measurements(meas, tr, want_spacepoints=True)
    for k in ("x", "y", "z", "r", "varZ", "varR"):
        assert np.array_equal(sp[k].view(np.uint32), got[k].view(np.uint32)), k
    assert _same_bits(seeds, orc.run(got))
    assert _same_bits(eng.run_measurements(meas, tr), seeds)
    eng.close()


from tests.fuzz import random_config as _random_config  # noqa: E402


def test_random_configurations_match_oracle(plugin, O):
    """Differential fuzzing: 48 random configurations (cuts, z binning, navigation, filter, seed confirmation,
    vertex windows) x one event each, bit-identical seeds in the reference's order."""
    from acts_b200 import events

    rng = np.random.default_rng(20260117)
    done = 0
    for trial in range(80):
        state = rng.bit_generator.state
        cfg, o = _random_config(rng, plugin.config_init)
        rng.bit_generator.state = state
        cfgo, _ = _random_config(rng, O.config_init)
        try:
            plugin.plan_info(cfg)
        except plugin.SeedingError:
            continue  # the reference constructor would throw as well (checked in test_plan_and_abi)
        ev = events.pileup_event(100 + trial, mu=float(rng.choice([3, 10, 25, 45])))
        zw = None
        if not o.get("seedConfirmation") and rng.integers(0, 4) == 0:
            zw = [(-60.0, -20.0), (5.0, 45.0)]
        eng = plugin.SeedingEngine(cfg)
        orc = O.Oracle(cfgo)
        got = eng.run(ev, z_windows=zw)  # no configuration is refused for its size: the spill class takes what shared memory cannot
        ref = orc.run(ev, z_windows=zw)
        assert _same_bits(got, ref), (trial, o)
        eng.close()
        done += 1
        if done == 48:
            break
    assert done == 48


ORTH_VARIANTS = [
    dict(),
    dict(interactionPointCut=1),
    dict(useExtraCuts=1),
    dict(useVariableMiddleSPRange=1, deltaRMiddleMinSPRange=25.0, deltaRMiddleMaxSPRange=40.0),
    dict(deltaPhiMax=0.025, maxSeedsPerSpM=4, sigmaScattering=2.0),
    dict(deltaPhiMax=0.2, zOutermostLayersMin=-650.0, zOutermostLayersMax=800.0, phiMin=-2.5, phiMax=2.0),
    dict(deltaRMinTop=float("nan"), deltaRMaxTop=float("nan"), deltaRMinBottom=6.0, deltaRMaxBottom=150.0),
    dict(deltaRMinTop=10.0, deltaRMaxTop=120.0, deltaRMinBottom=float("nan"), deltaRMaxBottom=float("nan")),
    dict(collisionRegionMin=-80.0, collisionRegionMax=120.0, cotThetaMax=3.0, deltaZMin=-300.0, deltaZMax=250.0),
    dict(useDeltaRinsteadOfTopRadius=1, compatSeedLimit=3, numSeedIncrement=100.0, impactWeightFactor=100.0),
]


@pytest.mark.parametrize("variant", range(len(ORTH_VARIANTS)))
def test_orthogonal_seeder_matches_oracle(plugin, O, variant):
    """OrthogonalTripletSeedingAlgorithm (k-d-tree candidate provider, both z-direction groups per middle) through
    b200seed_create_orthogonal: seeds bit-identical to the oracle IN THE REFERENCE'S ORDER (the oracle is pinned to
    the reference's own execute() on the same variants, tests/test_reference_pin.py)."""
    from acts_b200 import config, events

    over = ORTH_VARIANTS[variant]
    cfg, opt = config.orthogonal_config(plugin.orthogonal_config_init, **over)
    eng = plugin.SeedingEngine(cfg, orthogonal=opt)
    orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init, **over))
    evs = [events.muon_gun_event(variant), events.pileup_event(variant, mu=5), events.pileup_event(40 + variant, mu=30)]
    if variant in (0, 2, 4):
        evs.append(events.itk_pileup_event(variant, mu=10))
    refs = [orc.run(ev) for ev in evs]
    assert sum(r["bottom"].size for r in refs) > 0
    for k, (ev, ref) in enumerate(zip(evs, refs)):
        got = eng.run(ev)
        assert O.seed_set(got) == O.seed_set(ref), f"variant {variant} event {k}: seed set differs"
        assert _same_bits(got, ref), f"variant {variant} event {k}: seed order differs"
        cnt = eng.counters()
        assert cnt["nMiddles"] == ref["counters"]["nMiddles"]
        assert cnt["nBottomDoublets"] == ref["counters"]["nBottomDoublets"]
        assert cnt["nTopDoublets"] == ref["counters"]["nTopDoublets"]
    cols, offsets = events.concat_events(evs)
    for got, ref in zip(eng.run_batch(cols, offsets), refs):
        assert _same_bits(got, ref)
    eng.close()


def test_orthogonal_seeder_edge_cases_and_full_size(plugin, O):
    """Empty / tiny events (the tree degenerates to a leaf), quantised coordinates (equal keys in the tree's
    partition / sort and in the cotTheta sort), a <mu>=200 event, and seedConfirmation = true."""
    from acts_b200 import config, events

    cfg, opt = config.orthogonal_config(plugin.orthogonal_config_init)
    eng = plugin.SeedingEngine(cfg, orthogonal=opt)
    orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init))
    ev = events.pileup_event(0, mu=5)
    for n in (0, 1, 3, 4, 5, 9, 130):
        sub = {k: v[:n] for k, v in ev.items()}
        assert _same_bits(eng.run(sub), orc.run(sub)), n
    ev = events.pileup_event(5, mu=30)
    q = {k: (np.round(v * 4) / 4).astype(np.float32) if k in ("x", "y", "z") else v for k, v in ev.items()}
    q["r"] = (np.round(np.hypot(q["x"], q["y"]) * 2) / 2).astype(np.float32)
    ref = orc.run(q)
    assert ref["bottom"].size > 0 and _same_bits(eng.run(q), ref)
    big = events.pileup_event(3, mu=200)
    ref = orc.run(big)
    assert ref["bottom"].size > 50_000
    assert _same_bits(eng.run(big), ref)
    eng.close()
    # seedConfirmation through the k-d-tree seeder: one filter state (bestSeedQualityMap) across both z-direction
    # groups of a middle and across middles (.cpp:234-238) = the fixed-point replay over work items
    over = config.confirmation_overrides()
    cfg, opt = config.orthogonal_config(plugin.orthogonal_config_init, **over)
    eng = plugin.SeedingEngine(cfg, orthogonal=opt)
    orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init, **over))
    for ev in (events.pileup_event(5, mu=30), events.pileup_event(6, mu=60), events.itk_pileup_event(1, mu=20)):
        ref = orc.run(ev)
        assert ref["bottom"].size > 0
        assert _same_bits(eng.run(ev), ref)
    assert eng.counters()["nConfirmationRounds"] >= 2
    # grid-only entry points say so on an orthogonal handle
    for call in (lambda: eng.debug_grid(), lambda: eng.run(events.pileup_event(5, mu=5), z_windows=[(-10.0, 10.0)])):
        with pytest.raises(plugin.SeedingError) as exc:
            call()
        assert exc.value.code == config.ERR_UNSUPPORTED
    eng.close()


def test_orthogonal_tree_construction_corner_cases(plugin, O, monkeypatch):
    """The device replay of the reference's k-d tree construction (std::partition above 128 elements, std::sort
    below) on inputs that stress it: event sizes around the 4 / 128 thresholds, identical points, one azimuth for
    every point (degenerate phi splits: `pivot == end`, an empty right child), a selector that drops points
    (ordered compaction), empty events inside a batch, <mu>=100; and the host-layer builder as a cross-check."""
    from acts_b200 import config, events

    base = events.pileup_event(7, mu=30)
    n = base["x"].size
    cases = [{k: v[:m] for k, v in base.items()} for m in (126, 127, 128, 129, 130, 133, 257, 1025)]
    dup = {k: v.copy() for k, v in base.items()}
    for k in dup:  # 100 copies of one space point in the middle of a normal event (more than 128 identical points
        dup[k][1000:1100] = dup[k][999]  # make the reference itself recurse without end: KDTree.hpp:299-324)
    cases.append(dup)
    same_phi = {k: v.copy() for k, v in base.items()}
    ang = np.float32(0.7)
    same_phi["x"] = (same_phi["r"] * np.cos(ang)).astype(np.float32)  # one azimuth for everybody: the phi splits degenerate
    same_phi["y"] = (same_phi["r"] * np.sin(ang)).astype(np.float32)
    cases.append(same_phi)
    cases.append(events.pileup_event(2, mu=100))
    for over in (dict(), dict(useExtraCuts=1)):
        cfg, opt = config.orthogonal_config(plugin.orthogonal_config_init, **over)
        eng = plugin.SeedingEngine(cfg, orthogonal=opt)
        orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init, **over))
        refs = [orc.run(ev) for ev in cases]
        for i, (ev, ref) in enumerate(zip(cases, refs)):
            assert _same_bits(eng.run(ev), ref), (over, i)
        empty = {k: v[:0] for k, v in base.items()}
        batch = [cases[3], empty, cases[8], empty, cases[0]]
        cols, offsets = events.concat_events(batch)
        got = eng.run_batch(cols, offsets)
        for g, ev in zip(got, batch):
            assert _same_bits(g, orc.run(ev))
        eng.close()
    monkeypatch.setenv("B200SEED_KD_HOST", "1")
    cfg, opt = config.orthogonal_config(plugin.orthogonal_config_init)
    eng = plugin.SeedingEngine(cfg, orthogonal=opt)
    orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init))
    for ev in (cases[3], cases[8], cases[9]):
        assert _same_bits(eng.run(ev), orc.run(ev))
    eng.close()


@pytest.mark.parametrize("host_build", [False, True])
def test_orthogonal_refuses_the_input_the_reference_cannot_build(plugin, O, monkeypatch, host_build):
    """More than 128 space points with identical (phi, r, z) whose three mantissas are odd: the middle of the
    bounding box rounds up in every dimension, every point satisfies `x < mid`, the whole range becomes the left child
    again and the reference's constructor recurses without end (KDTree.hpp:299-324).  The engine reports it
    (INVALID_ARGUMENT) instead of looping; the same points with an even mantissa are peeled four at a time and work."""
    from acts_b200 import config, events

    if host_build:
        monkeypatch.setenv("B200SEED_KD_HOST", "1")
    ev = events.pileup_event(1, mu=5)
    x, y = np.float32(40.0), np.float32(0.0)
    for _ in range(10000):  # an azimuth with an odd mantissa (phi = atan2f(y, x) as the reference computes it)
        y = np.nextafter(y, np.float32(1e9))
        if int(np.float32(O.lib().oracle_atan2f(float(y), float(x))).view(np.uint32)) & 1:
            break
    odd = lambda v: (np.float32(v).view(np.uint32) | np.uint32(1)).view(np.float32)
    bad = {k: v.copy() for k, v in ev.items()}
    bad["x"][:200], bad["y"][:200] = x, y
    bad["z"][:200], bad["r"][:200] = odd(12.5), odd(40.0)
    cfg, opt = config.orthogonal_config(plugin.orthogonal_config_init)
    eng = plugin.SeedingEngine(cfg, orthogonal=opt)
    with pytest.raises(plugin.SeedingError) as exc:
        eng.run(bad)
    assert exc.value.code == config.ERR_INVALID_ARGUMENT
    assert eng.run(ev)["bottom"].size == O.Oracle(*config.orthogonal_config(O.orthogonal_config_init)).run(ev)["bottom"].size
    eng.close()


# ---------------------------------------------------------------------------
# Strip triplet path (TripletSeedFinder::Config::useStripInfo = true, TripletSeedFinder.cpp:164-406) through
# b200seed_run_strips; the oracle's restatement is pinned to the reference's own implementation in
# tests/test_reference_pin.py.
# ---------------------------------------------------------------------------
def _strip_event(kind, i, mu):
    from acts_b200 import events

    ev = dict(_event(kind, i, mu))
    ev["strip"] = events.strip_details(ev, seed=i)
    return ev


@pytest.mark.parametrize("name,kind,mu,ids,over", [
    ("seeding_py", "muon", 0, (0, 1), {}),
    ("pu200", "pileup", 5, (0, 1), {}),
    ("pu200", "pileup", 20, (0, 1, 2), {}),
    ("pu200", "pileup", 60, (3,), {}),
    ("pu200", "pileup", 20, (4, 5), dict(seedConfirmation=1)),
    ("pu200", "pileup", 20, (6,), dict(toleranceParam=0.6, interactionPointCut=1)),
    ("itk_like", "pileup", 20, (7,), dict(toleranceParam=3.0, useDeltaRinsteadOfTopRadius=1)),
    ("itk_conf", "pileup", 20, (8,), {}),
])
def test_strip_triplet_path_matches_oracle(plugin, O, name, kind, mu, ids, over):
    eng = plugin.SeedingEngine(make_config(name, plugin.config_init).update(**over))
    orc = O.Oracle(make_config(name, O.config_init).update(**over))
    total = 0
    for i in ids:
        ev = _strip_event(kind, i, mu)
        for diff in (float("inf"), 0.4, 0.06, 0.0):
            if mu >= 60 and diff == float("inf"):
                continue  # (every bottom x top pair of every middle: minutes on the CPU side)
            got = eng.run(ev, strip_cot_theta_diff_max=diff)
            ref = orc.run(ev, strip_cot_theta_diff_max=diff)
            assert O.seed_set(got) == O.seed_set(ref), f"{name} event {i} cotThetaDiffMax {diff}: seed set differs"
            assert _same_bits(got, ref), f"{name} event {i} cotThetaDiffMax {diff}: order differs"
            cnt = eng.counters()
            assert cnt["nTripletTests"] == ref["counters"]["nTripletTests"]
            assert cnt["nCandidates"] == ref["counters"]["nCandidates"]
            total += ref["bottom"].size
        # the pixel path on the same handle afterwards is untouched by the strip call
        assert _same_bits(eng.run(ev), orc.run(ev))
    assert total > 0
    eng.close()


def test_strip_triplet_path_edge_cases(plugin, O):
    """Empty event, three points, parallel strips, quantised coordinates (cotTheta ties), and the argument checks
    of the entry point.  (All-zero details give 0 / 0 = NaN positions, curvatures and weights: the reference then
    sorts and heaps NaNs, whose order is whatever libstdc++'s comparisons happen to leave -- the oracle reproduces
    that on the CPU (tests/test_reference_pin.py), the device is only required not to fail on it.)"""
    from acts_b200 import events

    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    orc = O.Oracle(make_config("pu200", O.config_init))
    ev = _strip_event("pileup", 11, 10)
    empty = {k: np.zeros(0, np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")}
    empty["strip"] = np.zeros((0, 12), np.float32)
    three = {k: (v[:3] if k != "strip" else v[:3]) for k, v in ev.items()}
    zero = dict(ev, strip=np.zeros_like(ev["strip"]))
    par = dict(ev, strip=ev["strip"].copy())
    par["strip"][:, 9:12] = par["strip"][:, 6:9]
    q = {k: (np.round(ev[k] * 4) / 4).astype(np.float32) if k in ("x", "y", "z") else ev[k] for k in ev}
    q["r"] = np.hypot(q["x"].astype(np.float64), q["y"].astype(np.float64)).astype(np.float32)
    for case in (empty, three, par, q):
        for diff in (float("inf"), 0.1):
            got = eng.run(case, strip_cot_theta_diff_max=diff)
            ref = orc.run(case, strip_cot_theta_diff_max=diff)
            assert _same_bits(got, ref)
    got = eng.run(zero, strip_cot_theta_diff_max=0.1)
    assert got["bottom"].size == orc.run(zero, strip_cot_theta_diff_max=0.1)["bottom"].size
    with pytest.raises(plugin.SeedingError) as ei:
        eng.run(ev, strip_cot_theta_diff_max=float("nan"))
    assert ei.value.code != 0
    eng.close()
    from acts_b200 import config

    ocfg, oopt = config.orthogonal_config(plugin.orthogonal_config_init)
    oeng = plugin.SeedingEngine(ocfg, orthogonal=oopt)
    with pytest.raises(plugin.SeedingError):
        oeng.run(ev, strip_cot_theta_diff_max=1.0)
    oeng.close()


ITK_STRIP_FILTER = dict(impactWeightFactor=1.0, compatSeedLimit=4, numSeedIncrement=1.0, seedWeightIncrement=10100.0,
                        maxSeedsPerSpMConf=100, maxQualitySeedsPerSpMConf=100, maxSeedsPerSpM=4)


@pytest.mark.parametrize("conf", [False, True])
def test_itk_strip_filter_block_collectors_of_100(plugin, O, conf):
    """The filter block of the ITk STRIP configuration (Python/Examples/python/itk.py:499-506: collector capacities
    100 / 100 with maxSeedsPerSpM = 4), with and without seedConfirmation, on smeared and on lattice-quantised
    events (equal weights inside the collector: the literal heap replay runs with 100-entry heaps), through the
    pixel and the strip triplet path."""
    from acts_b200 import config as cm
    from acts_b200 import events

    extra = dict(cm.confirmation_overrides(), **ITK_STRIP_FILTER) if conf else dict(ITK_STRIP_FILTER)
    eng = plugin.SeedingEngine(make_config("pu200", plugin.config_init).update(**extra))
    orc = O.Oracle(make_config("pu200", O.config_init).update(**extra))
    ties = 0
    for i, mu, step in ((0, 30, 0.0), (1, 60, 0.0), (2, 30, 0.5), (3, 60, 2.0)):
        ev = dict(events.pileup_event(i, mu=mu))
        if step > 0:
            for k in ("x", "y", "z"):
                ev[k] = (np.round(ev[k] / np.float32(step)) * np.float32(step)).astype(np.float32)
            ev["r"] = np.sqrt(ev["x"].astype(np.float64) ** 2 + ev["y"].astype(np.float64) ** 2).astype(np.float32)
            ev["varZ"] = np.full_like(ev["varZ"], 0.01)
            ev["varR"] = np.full_like(ev["varR"], 0.01)
        got, ref = eng.run(ev), orc.run(ev)
        assert ref["quality"].size > 100
        ties += ref["counters"]["nWeightTieMiddles"]
        assert _same_bits(got, ref), (conf, i, "pixel path")
        ev["strip"] = events.strip_details(ev, seed=i)
        got, ref = eng.run(ev, strip_cot_theta_diff_max=0.2), orc.run(ev, strip_cot_theta_diff_max=0.2)
        assert _same_bits(got, ref), (conf, i, "strip path")
    assert ties > 0
    eng.close()


@pytest.mark.parametrize("name,mu,ids", [("itk_strip", 60, (0, 1)), ("itk_strip", 200, (2,)), ("itk_strip_grid", 200, (3,))])
def test_verbatim_itk_strip_configuration(plugin, O, name, mu, ids):
    """itkSeedingAlgConfig(StripSpacePoints) verbatim (Python/Examples/python/itk.py:458-506: rMax 1200 mm, deltaR up
    to 600 mm, deltaZMax 900 mm, seedConfirmation with collectors of 100 / 100, seedWeightIncrement 10100) on
    ITk-strip-shaped events: the pixel triplet path (what the reference's algorithm runs with this configuration)
    and the strip triplet path on the same points."""
    from acts_b200 import events

    eng = plugin.SeedingEngine(make_config(name, plugin.config_init))
    orc = O.Oracle(make_config(name, O.config_init))
    for i in ids:
        ev = events.itk_strip_event(i, mu=mu)
        got, ref = eng.run(ev), orc.run(ev)
        assert ref["quality"].size > 1000
        assert _same_bits(got, ref), (name, i, "pixel path")
        got, ref = eng.run(ev, strip_cot_theta_diff_max=0.1), orc.run(ev, strip_cot_theta_diff_max=0.1)
        assert ref["quality"].size > 500
        assert _same_bits(got, ref), (name, i, "strip path")
    eng.close()


@pytest.mark.parametrize("mb", ["0", "1"])
def test_orthogonal_hit_list_arena_off_or_full(plugin, O, monkeypatch, mb):
    """The count pass of the orthogonal seeder hands the survivor positions of every tree walk to the fill pass
    (B200SEED_KD_HIT_MB, default 4096).  Off (0) and nearly always full (1 MB): a walk whose list did not fit is
    searched again by the fill pass -- same seeds in the same order."""
    from acts_b200 import config, events

    monkeypatch.setenv("B200SEED_KD_HIT_MB", mb)
    ocfg, oopt = config.orthogonal_config(plugin.orthogonal_config_init)
    eng = plugin.SeedingEngine(ocfg, orthogonal=oopt)
    orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init))
    evs = [events.pileup_event(i, mu=m) for i, m in ((0, 30), (1, 60))]
    for ev in evs:
        assert _same_bits(eng.run(ev), orc.run(ev))
    cols, offsets = events.concat_events(evs)
    for got, ev in zip(eng.run_batch(cols, offsets), evs):
        assert _same_bits(got, orc.run(ev))
    eng.close()
