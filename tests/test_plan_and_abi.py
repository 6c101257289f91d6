"""Host-side product code that needs no GPU: the C ABI library loads, exports every
declared symbol, validates configurations like the reference constructors and
derives the same constants as the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import CONFIGS, make_config


@pytest.fixture(scope="module")
def plugin(built):
    from acts_b200 import plugin

    return plugin


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle

    return oracle


def test_library_exports_every_declared_symbol(plugin):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "acts_b200_seeding.h")).read()
    declared = set(re.findall(r"\b(b200seed_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(plugin.EXPORTED_SYMBOLS)
    L = plugin.lib()
    for name in sorted(declared):
        assert getattr(L, name) is not None, name


def test_defaults_equal_reference_defaults(plugin, O):
    from acts_b200.config import Config

    a, b = Config(), Config()
    plugin.config_init(C.byref(a))
    O.config_init(C.byref(b))
    assert bytes(a) == bytes(b)
    assert a.struct_size == C.sizeof(Config)
    assert a.maxSeedsPerSpM == 5 and a.compatSeedLimit == 2 and a.deltaRMin == 5.0 and a.rMax == 600.0
    assert np.isnan(a.deltaRMinTop) and np.isinf(a.numSeedIncrement)


@pytest.mark.parametrize("name", CONFIGS)
def test_plan_matches_oracle_derivation(plugin, O, name):
    pi = plugin.plan_info(make_config(name, plugin.config_init))
    oi = O.Oracle(make_config(name, O.config_init)).info()
    for f in ("phiBins", "zBins", "rBins", "nGlobalBins", "minHelixDiameter2", "highland", "sigmapT2perRadius",
              "multipleScattering2", "deltaRMinBottom", "deltaRMaxBottom", "deltaRMinTop", "deltaRMaxTop"):
        assert getattr(pi, f) == getattr(oi, f), f


def test_plan_matches_oracle_on_random_cut_sets(plugin, O):
    """The host plan (phi bin count of SpacePointGridPhiBinning.cpp:20-95, derived finder constants) against the
    oracle's derivation over random minPt / bField / rMax / deltaRMax / impactMax / coverage -- including the
    combinations both refuse, with the same status code."""
    from acts_b200 import config as cm

    rng = np.random.default_rng(424242)
    accepted = refused = 0
    for _ in range(400):
        over = dict(minPt=float(rng.choice([0.1, 0.4, 0.5, 0.9, 2.0, 10.0])), bFieldInZ=float(rng.choice([0.0, 1.0, 2.0, 3.8])) * cm.T,
                    rMax=float(rng.choice([100.0, 200.0, 320.0, 1200.0])), deltaRMax=float(rng.choice([60.0, 150.0, 280.0, 600.0])),
                    impactMax=float(rng.choice([0.5, 3.0, 20.0, 150.0])), phiBinDeflectionCoverage=int(rng.choice([1, 2, 3, 5])),
                    maxPhiBins=int(rng.choice([20, 200, 10000])), sigmaScattering=float(rng.choice([2.0, 5.0])),
                    radLengthPerSeed=float(rng.choice([0.05, 0.1])))
        try:
            want = O.Oracle(make_config("pu200", O.config_init).update(**over)).info()
        except O.OracleError as e:
            with pytest.raises(plugin.SeedingError) as ei:
                plugin.plan_info(make_config("pu200", plugin.config_init).update(**over))
            assert ei.value.code == e.code, over
            refused += 1
            continue
        got = plugin.plan_info(make_config("pu200", plugin.config_init).update(**over))
        for f in ("phiBins", "zBins", "rBins", "nGlobalBins"):
            assert getattr(got, f) == getattr(want, f), (f, over)
        for f in ("minHelixDiameter2", "highland", "sigmapT2perRadius", "multipleScattering2"):
            a, b = np.float32(getattr(got, f)), np.float32(getattr(want, f))
            assert a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b)), (f, over)
        accepted += 1
    assert accepted > 100 and refused > 10


@pytest.mark.parametrize("name", CONFIGS)
def test_neighbour_tables_match_oracle(plugin, O, name):
    cfg = make_config(name, plugin.config_init)
    t = plugin.plan_tables(cfg)
    orc = O.Oracle(make_config(name, O.config_init))
    info = orc.info()
    nz, nr = info.zBins, info.rBins
    for g, gbin in enumerate(t["navBins"]):
        r_loc = int(gbin) % (nr + 2)
        z_loc = (int(gbin) // (nr + 2)) % (nz + 2)
        p_loc = int(gbin) // ((nr + 2) * (nz + 2))
        for top, offs, bins in ((False, t["botOffsets"], t["botBins"]), (True, t["topOffsets"], t["topBins"])):
            ref = orc.find_bins(p_loc, z_loc, r_loc, top)
            # never-fillable under/overflow bins are dropped by the plan (host_plan.cpp)
            def fillable(b):
                rl = b % (nr + 2); zl = (b // (nr + 2)) % (nz + 2)
                return 1 <= rl <= nr and 1 <= zl <= nz
            ref = [int(b) for b in ref if fillable(int(b))]
            assert bins[offs[g]:offs[g + 1]].tolist() == ref


def test_config_errors_map_to_reference_exceptions(plugin, O):
    from acts_b200 import config as cm

    cases = [
        (dict(minPt=0.010), cm.ERR_DOMAIN),                                      # std::domain_error
        (dict(phiMin=-4.0), cm.ERR_RUNTIME),                                     # std::runtime_error
        (dict(zMin=100.0, zMax=-100.0), cm.ERR_RUNTIME),
        (dict(zBinEdges=[-1.0, 0.0, 1.0], zBinsCustomLooping=[1, 3]), cm.ERR_INVALID_ARGUMENT),
        (dict(zBinEdges=[-2000.0, 0.0, 2000.0], zBinsCustomLooping=[1, 1]), cm.ERR_INVALID_ARGUMENT),
        (dict(phiMin=1.0, phiMax=1.0), cm.ERR_INVALID_ARGUMENT),
    ]
    for override, code in cases:
        with pytest.raises(plugin.SeedingError) as e1:
            plugin.plan_info(make_config("pu200", plugin.config_init).update(**override))
        assert e1.value.code == code, override
        with pytest.raises(O.OracleError) as e2:
            O.Oracle(make_config("pu200", O.config_init).update(**override))
        assert e2.value.code == code, override


def test_unsupported_configs_are_rejected_loudly(plugin):
    from acts_b200 import config as cm

    # collector capacities up to 128 are fine (itk.py:504-505 uses 100); a middle can return at most 16 seeds
    for ok in (dict(maxSeedsPerSpMConf=100), dict(seedConfirmation=1, maxSeedsPerSpMConf=100, maxQualitySeedsPerSpMConf=100)):
        plugin.plan_info(make_config("pu200", plugin.config_init).update(**ok))
    for override in (dict(seedConfirmation=1, maxQualitySeedsPerSpMConf=129), dict(compatSeedLimit=9), dict(maxSeedsPerSpMConf=129),
                     dict(maxSeedsPerSpMConf=17, maxSeedsPerSpM=16)):
        with pytest.raises(plugin.SeedingError) as ei:
            plugin.plan_info(make_config("pu200", plugin.config_init).update(**override))
        assert ei.value.code == cm.ERR_UNSUPPORTED
    bad = make_config("pu200", plugin.config_init)
    bad.struct_size = 12
    with pytest.raises(plugin.SeedingError) as ei:
        plugin.plan_info(bad)
    assert ei.value.code == cm.ERR_INVALID_ARGUMENT


def test_no_cpu_fallback_without_a_device(plugin):
    """Without a CUDA device b200seed_create must fail with ERR_CUDA, never compute on the host."""
    import torch

    from acts_b200 import config as cm

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(plugin.SeedingError) as ei:
        plugin.SeedingEngine(make_config("pu200", plugin.config_init))
    assert ei.value.code == cm.ERR_CUDA


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under acts_b200/ may import, include or load it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s+[\"<][^\">]*oracle)|(libseeding_oracle)|(oracle/)", re.M)
    for dirpath, _, files in os.walk(os.path.join(root, "acts_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert pat.search(text) is None, f
