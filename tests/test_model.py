"""The sequential CPU model of the GPU algorithm (tests/model) equals the oracle.

This is where the exactness of the restructuring is established without a GPU:
binary-searched r windows, (cotTheta, emission index) order + pruned libstdc++
introsort replay for ties, per-bottom independent top windows (prefix-max
formulation), independent filter weights, replayed bounded heap."""
import numpy as np
import pytest

from tests.conftest import CONFIGS, make_config

KEYS = ("bottom", "middle", "top", "quality", "vertexZ")


@pytest.fixture(scope="module")
def env(built):
    from acts_b200 import events, plugin
    from oracle import oracle as O
    from tests.model import model as M

    return events, plugin, O, M


def test_libstdcxx_sort_replay_matches_std_sort(env):
    M = env[3]
    L = M.lib()
    assert L.model_check_std_sort(1, 3000, 2000, 7) == 0
    assert L.model_check_std_sort(2, 40, 20000, 3) == 0
    assert L.model_check_std_sort(3, 100000, 10, 50) == 0
    assert L.model_check_std_sort_killer(4096) == 0   # forces the heap-sort fallback


def test_pruned_tie_replay_matches_std_sort(env):
    L = env[3].lib()
    assert L.model_check_tie_replay(1, 3000, 2000, 3) == 0
    assert L.model_check_tie_replay(2, 60000, 50, 6) == 0
    assert L.model_check_tie_replay(3, 40, 20000, 2) == 0
    assert L.model_check_tie_replay(4, 2000, 2000, 40) == 0


def test_warp_parallel_replay_formulation_matches_std_sort(env):
    """Scalar emulation of the device's warp-parallel partition (32 + 32 elements per step, k-th
    misplaced left swapped with k-th misplaced right) against std::sort."""
    L = env[3].lib()
    assert L.model_check_warp_replay(1, 3000, 2000, 3, 0) == 0
    assert L.model_check_warp_replay(2, 60000, 60, 6, 0) == 0
    assert L.model_check_warp_replay(3, 200, 20000, 2, 0) == 0
    assert L.model_check_warp_replay(4, 2500, 2000, 40, 0) == 0
    assert L.model_check_warp_replay(5, 1000, 2000, 500, 0) == 0


def test_heap_replay_matches_libstdcxx(env):
    assert env[3].lib().model_check_heap(5, 200000) == 0


def test_atan2f_replay_matches_host_libm(env):
    """glibc's atan2f (what the reference calls, GridTripletSeedingAlgorithm.cpp:219) op for op."""
    assert env[3].lib().model_check_atan2f(99, 30_000_000) == 0


@pytest.mark.parametrize("name", CONFIGS)
def test_model_equals_oracle(env, name):
    events, plugin, O, M = env
    cfg = make_config(name, plugin.config_init)
    orc = O.Oracle(make_config(name, O.config_init))
    for i, mu in ((0, 10), (2, 40)):
        ev = events.muon_gun_event(i) if name == "seeding_py" else events.pileup_event(i, mu=mu)
        ref = orc.run(ev, want_grid=True)
        for tie_mode in (1, 2):   # full std::sort replay, canonical + pruned replay (what the kernel does)
            got = M.run(cfg, ref["grid"], tie_mode=tie_mode)
            for k in KEYS:
                assert np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32)), (name, i, tie_mode, k)
            assert got["stats"]["nCandidates"] == ref["counters"]["nCandidates"]
        stable = orc.run(ev, sort_mode=O.Oracle.STABLE, want_grid=True)
        got0 = M.run(cfg, stable["grid"], tie_mode=0)
        for k in KEYS:
            assert np.array_equal(got0[k].view(np.uint32), stable[k].view(np.uint32))


def test_seed_confirmation_fixed_point_equals_sequential_map(env):
    """seedConfirmation: records + collector replay + fixed-point rounds (what k_conf_replay and the round loop do)
    reproduce the reference's sequential bestSeedQualityMap pass; the rounds really iterate (> 2)."""
    events, plugin, O, M = env
    from acts_b200 import config as cm

    deepest = 0
    for base, extra in (("itk_conf", {}), ("pu200", cm.confirmation_overrides()),
                        ("pu200", dict(cm.confirmation_overrides(), maxQualitySeedsPerSpMConf=0)),
                        ("pu200", dict(cm.confirmation_overrides(), maxSeedsPerSpM=0, maxSeedsPerSpMConf=1))):
        cfg = make_config(base, plugin.config_init).update(**extra)
        orc = O.Oracle(make_config(base, O.config_init).update(**extra))
        for i, mu in ((1, 30), (4, 60)):
            ev = events.pileup_event(i, mu=mu)
            ref = orc.run(ev, want_grid=True)
            got = M.run(cfg, ref["grid"], tie_mode=2)
            assert ref["quality"].size > 100
            for k in KEYS:
                assert np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32)), (base, extra, i, k)
            deepest = max(deepest, got["stats"]["nConfirmationRounds"])
    assert deepest > 2


def test_model_bin_index_equals_oracle(env):
    events, plugin, O, M = env
    for name in ("pu200", "itk_like"):
        cfg = make_config(name, plugin.config_init)
        orc = O.Oracle(make_config(name, O.config_init))
        ev = events.pileup_event(7, mu=10)
        n = 4000
        got = M.bin_index(cfg, ev["x"][:n], ev["y"][:n], ev["z"][:n], ev["r"][:n])
        ref = np.array([orc.bin_index(float(O.lib().oracle_atan2f(float(y), float(x))), float(z), float(r))
                        for x, y, z, r in zip(ev["x"][:n], ev["y"][:n], ev["z"][:n], ev["r"][:n])])
        assert np.array_equal(got, ref)


def test_branch_free_pair_classification_equals_the_early_exit_form():
    """classify_pair_flat (device scans) evaluates every term of the triplet test unconditionally and selects the
    class at the end: same class as the reference's early-exit order on 2e7 random pairs incl. dU == 0."""
    from tests.model import model

    L = model.lib()
    L.model_check_flat_classify.restype = __import__("ctypes").c_int64
    L.model_check_flat_classify.argtypes = [__import__("ctypes").c_uint64, __import__("ctypes").c_int64]
    assert L.model_check_flat_classify(7, 20_000_000) == 0


def test_partition_rank_pairing_equals_std_partition(env):
    """k_kd_split (orthogonal seeder, k-d tree construction on the device) replays libstdc++'s bidirectional
    std::partition by ranks: k-th misplaced element from the left <-> k-th misplaced element from the right."""
    L = env[3].lib()
    L.model_check_partition_pairing.restype = __import__("ctypes").c_int64
    L.model_check_partition_pairing.argtypes = [__import__("ctypes").c_uint64, __import__("ctypes").c_int, __import__("ctypes").c_int]
    assert L.model_check_partition_pairing(1, 40, 20000) == 0
    assert L.model_check_partition_pairing(2, 3000, 2000) == 0
    assert L.model_check_partition_pairing(3, 200000, 20) == 0


def test_strip_window_per_bottom_equals_the_sequential_subrange_walk(env):
    """Strip triplet path: the device finds every bottom's cot(theta) window on its own (binary search + scan); the
    reference advances one shared subrange through the bottoms (TripletSeedFinder.cpp:226-238,397-403).  Same pairs,
    same order -- random sorted lists with ties, cotThetaDiffMax = inf / 0 / finite."""
    import ctypes as C

    L = env[3].lib()
    L.model_check_strip_window.restype = C.c_int64
    L.model_check_strip_window.argtypes = [C.c_uint64, C.c_int, C.c_int]
    assert L.model_check_strip_window(1, 12, 200000) == 0
    assert L.model_check_strip_window(2, 300, 3000) == 0


@pytest.mark.parametrize("name,over", [("pu200", {}), ("seeding_py", {}), ("pu200", dict(toleranceParam=0.6)),
                                       ("itk_conf", {})])
def test_model_strip_triplet_path_equals_oracle(env, name, over):
    """The device formulation of the strip triplet path -- per-bottom cot(theta) windows, eval_strip_pair /
    strip_calibrate / strip_derive of seed_math.h compiled for the host -- against the oracle's restatement of
    createStripTripletTopCandidates (itself pinned to the reference, tests/test_reference_pin.py)."""
    events, plugin, O, M = env
    cfg = make_config(name, plugin.config_init).update(**over)
    orc = O.Oracle(make_config(name, O.config_init).update(**over))
    for i, mu in ((0, 10), (2, 30)):
        ev = dict(events.muon_gun_event(i) if name == "seeding_py" else events.pileup_event(i, mu=mu))
        ev["strip"] = events.strip_details(ev, seed=i)
        grid = orc.run(ev, want_grid=True)["grid"]
        for diff in (float("inf"), 0.3, 0.05):
            ref = orc.run(ev, strip_cot_theta_diff_max=diff)
            got = M.run(cfg, grid, tie_mode=2, strip=ev["strip"], cot_theta_diff_max=diff)
            for k in KEYS:
                assert np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32)), (name, i, diff, k)
            assert got["stats"]["nTripletTests"] == ref["counters"]["nTripletTests"]
            assert got["stats"]["nCandidates"] == ref["counters"]["nCandidates"]


def test_top_k_selection_claim_against_the_literal_collector(env):
    """k_seed_middles keeps min(nLow, maxSeedsPerSpM + 1) + 1 weights and trusts them when they differ: the literal
    bounded heap of any capacity nLow then returns the same first entries (random weights, heavy ties, nLow up to 128)."""
    import ctypes as C

    L = env[3].lib()
    L.model_check_topk_claim.restype = C.c_int64
    L.model_check_topk_claim.argtypes = [C.c_uint64, C.c_int]
    assert L.model_check_topk_claim(5, 200000) == 0
