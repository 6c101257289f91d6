"""Random but valid configurations inside the supported envelope (differential fuzzing), shared by the
CPU pin tests (oracle vs the unmodified reference) and the GPU parity tests (CUDA vs oracle)."""
from tests.conftest import make_config


def random_config(rng, init):
    """A random but valid configuration inside the supported envelope (differential fuzzing)."""
    from acts_b200 import config as cm

    cfg = make_config("pu200", init)
    o = {}
    o["sigmaScattering"] = float(rng.choice([2.0, 5.0, 20.0, 50.0]))
    o["minPt"] = float(rng.choice([0.4, 0.5, 0.9, 2.0]))
    o["impactMax"] = float(rng.choice([1.0, 3.0, 10.0, 20.0]))
    o["cotThetaMax"] = float(rng.choice([3.0, 7.40627, 10.0]))
    o["deltaRMinTop"], o["deltaRMaxTop"] = float(rng.choice([1.0, 5.0, 10.0])), float(rng.choice([80.0, 150.0, 300.0]))
    o["deltaRMinBottom"], o["deltaRMaxBottom"] = float(rng.choice([1.0, 5.0, 10.0])), float(rng.choice([80.0, 150.0, 300.0]))
    o["collisionRegionMin"], o["collisionRegionMax"] = float(rng.choice([-250.0, -150.0, -50.0])), float(rng.choice([60.0, 150.0, 250.0]))
    o["maxSeedsPerSpM"] = int(rng.choice([0, 1, 2, 5]))
    o["maxSeedsPerSpMConf"] = int(rng.choice([1, 3, 5, 16, 40, 100]))
    o["compatSeedLimit"] = int(rng.choice([0, 1, 2, 3, 8]))
    o["compatSeedWeight"] = float(rng.choice([100.0, 200.0]))
    o["impactWeightFactor"] = float(rng.choice([1.0, 100.0]))
    o["deltaInvHelixDiameter"] = float(rng.choice([0.00003, 0.0003]))
    o["deltaRMin"] = float(rng.choice([1.0, 5.0, 20.0]))
    o["numPhiNeighbors"] = int(rng.choice([0, 1, 2]))
    o["phiBinDeflectionCoverage"] = int(rng.choice([1, 2, 3]))
    o["interactionPointCut"] = int(rng.integers(0, 2))
    o["useExtraCuts"] = int(rng.integers(0, 2))
    o["useDeltaRinsteadOfTopRadius"] = int(rng.integers(0, 2))
    o["helixCutTolerance"] = float(rng.choice([1.0, 2.0]))
    o["toleranceParam"] = float(rng.choice([1.1, 2.0]))
    o["radLengthPerSeed"] = float(rng.choice([0.05, 0.1]))
    if rng.integers(0, 2):
        o["deltaZMin"], o["deltaZMax"] = float(rng.choice([-300.0, -80.0])), float(rng.choice([120.0, 400.0]))
    if rng.integers(0, 3) == 0:
        o["seedWeightIncrement"], o["numSeedIncrement"] = 7.5, float(rng.choice([0.0, 1.0]))
    zmode = int(rng.integers(0, 3))
    if zmode == 1:
        o["zBinEdges"] = [-2000.0, -300.0, 0.0, 300.0, 2000.0]
    elif zmode == 2:
        o["zBinEdges"] = [-2000.0, -800.0, -250.0, 250.0, 800.0, 2000.0]
        o["zBinNeighborsTop"] = [(0, 0), (-1, 0), (-1, 1), (0, 1), (0, 0)]
        o["zBinNeighborsBottom"] = [(0, 1), (0, 1), (-1, 1), (-1, 0), (-1, 0)]
        o["zBinsCustomLooping"] = [3, 2, 4, 1, 5]
        if rng.integers(0, 2):
            o["rRangeMiddleSP"] = [(60.0, 130.0), (50.0, 120.0), (40.0, 180.0), (50.0, 120.0), (60.0, 130.0)]
    if rng.integers(0, 3) == 0:
        o["useVariableMiddleSPRange"] = 1
        o["deltaRMiddleMinSPRange"], o["deltaRMiddleMaxSPRange"] = 10.0, float(rng.choice([10.0, 40.0]))
    if rng.integers(0, 2):
        conf = cm.confirmation_overrides()
        conf["maxQualitySeedsPerSpMConf"] = int(rng.choice([0, 1, 5, 100]))
        conf["zOriginWeightFactor"] = float(rng.choice([0.0, 1.0]))
        conf["centralSeedConfirmationRange"]["nTopForSmallR"] = int(rng.choice([1, 2, 3]))
        conf["forwardSeedConfirmationRange"]["rMaxSeedConf"] = float(rng.choice([80.0, 140.0]))
        for k in ("compatSeedLimit", "maxSeedsPerSpMConf", "compatSeedWeight", "impactWeightFactor"):
            conf.pop(k)
        o.update(conf)
    return cfg.update(**o), o
