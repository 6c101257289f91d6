"""ctypes mirror of ``include/acts_b200_seeding.h`` and the canonical configurations.

The field names are the reference's ``GridTripletSeedingAlgorithm::Config``
(Examples/Algorithms/TrackFinding/include/ActsExamples/TrackFinding/
GridTripletSeedingAlgorithm.hpp:34-244).  This module is pure host-side
marshalling; it never computes anything the engine computes.
"""
from __future__ import annotations

import ctypes as C
import math

ABI_VERSION = 2

# status codes (acts_b200_seeding.h)
OK, ERR_INVALID_ARGUMENT, ERR_RUNTIME, ERR_DOMAIN, ERR_UNSUPPORTED, ERR_CUDA, ERR_CAPACITY, ERR_OVERFLOW = range(8)

# Acts::UnitConstants (Core/include/Acts/Definitions/Units.hpp:85,143,149,168)
mm = 1.0
GeV = 1.0
MeV = 1e-3
T = 0.000299792458


class SeedConfirmationRange(C.Structure):
    _fields_ = [
        ("zMinSeedConf", C.c_float),
        ("zMaxSeedConf", C.c_float),
        ("rMaxSeedConf", C.c_float),
        ("nTopForLargeR", C.c_uint64),
        ("nTopForSmallR", C.c_uint64),
        ("seedConfMinBottomRadius", C.c_float),
        ("seedConfMaxZOrigin", C.c_float),
        ("minImpactSeedConf", C.c_float),
    ]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("struct_size", C.c_uint32),
        ("bFieldInZ", C.c_float),
        ("minPt", C.c_float),
        ("cotThetaMax", C.c_float),
        ("impactMax", C.c_float),
        ("deltaRMin", C.c_float),
        ("deltaRMax", C.c_float),
        ("deltaRMinTop", C.c_float),
        ("deltaRMaxTop", C.c_float),
        ("deltaRMinBottom", C.c_float),
        ("deltaRMaxBottom", C.c_float),
        ("rMin", C.c_float),
        ("rMax", C.c_float),
        ("zMin", C.c_float),
        ("zMax", C.c_float),
        ("phiMin", C.c_float),
        ("phiMax", C.c_float),
        ("phiBinDeflectionCoverage", C.c_int32),
        ("maxPhiBins", C.c_int32),
        ("zBinNeighborsTop", C.POINTER(C.c_int32)),
        ("nZBinNeighborsTop", C.c_uint32),
        ("zBinNeighborsBottom", C.POINTER(C.c_int32)),
        ("nZBinNeighborsBottom", C.c_uint32),
        ("numPhiNeighbors", C.c_int32),
        ("zBinEdges", C.POINTER(C.c_float)),
        ("nZBinEdges", C.c_uint32),
        ("zBinsCustomLooping", C.POINTER(C.c_uint64)),
        ("nZBinsCustomLooping", C.c_uint32),
        ("rMinMiddle", C.c_float),
        ("rMaxMiddle", C.c_float),
        ("useVariableMiddleSPRange", C.c_uint8),
        ("rRangeMiddleSP", C.POINTER(C.c_float)),
        ("nRRangeMiddleSP", C.c_uint32),
        ("deltaRMiddleMinSPRange", C.c_float),
        ("deltaRMiddleMaxSPRange", C.c_float),
        ("deltaZMin", C.c_float),
        ("deltaZMax", C.c_float),
        ("interactionPointCut", C.c_uint8),
        ("collisionRegionMin", C.c_float),
        ("collisionRegionMax", C.c_float),
        ("helixCutTolerance", C.c_float),
        ("sigmaScattering", C.c_float),
        ("radLengthPerSeed", C.c_float),
        ("toleranceParam", C.c_float),
        ("deltaInvHelixDiameter", C.c_float),
        ("compatSeedWeight", C.c_float),
        ("impactWeightFactor", C.c_float),
        ("zOriginWeightFactor", C.c_float),
        ("maxSeedsPerSpM", C.c_uint32),
        ("compatSeedLimit", C.c_uint64),
        ("seedWeightIncrement", C.c_float),
        ("numSeedIncrement", C.c_float),
        ("seedConfirmation", C.c_uint8),
        ("centralSeedConfirmationRange", SeedConfirmationRange),
        ("forwardSeedConfirmationRange", SeedConfirmationRange),
        ("maxSeedsPerSpMConf", C.c_uint32),
        ("maxQualitySeedsPerSpMConf", C.c_uint32),
        ("useDeltaRinsteadOfTopRadius", C.c_uint8),
        ("useExtraCuts", C.c_uint8),
        ("useVertexZCuts", C.c_uint8),
        ("vertexZNSigma", C.c_double),
        ("vertexZMargin", C.c_double),
        ("relaxedFloat", C.c_uint8),
    ]

    # python-side keep-alives for the (pointer, count) members
    _ARRAYS = {
        "zBinNeighborsTop": (C.c_int32, "nZBinNeighborsTop", 2),
        "zBinNeighborsBottom": (C.c_int32, "nZBinNeighborsBottom", 2),
        "zBinEdges": (C.c_float, "nZBinEdges", 1),
        "zBinsCustomLooping": (C.c_uint64, "nZBinsCustomLooping", 1),
        "rRangeMiddleSP": (C.c_float, "nRRangeMiddleSP", 2),
    }

    def set_array(self, name, values):
        """Set a vector member; ``values`` is a flat list or a list of pairs."""
        ctype, count_name, width = self._ARRAYS[name]
        flat = []
        for v in values:
            if width == 2:
                flat.extend(v)
            else:
                flat.append(v)
        if not hasattr(self, "_keep"):
            self._keep = {}
        arr = (ctype * max(len(flat), 1))(*flat)
        self._keep[name] = arr
        setattr(self, name, C.cast(arr, C.POINTER(ctype)))
        setattr(self, count_name, len(flat) // width)

    def update(self, **kw):
        for k, v in kw.items():
            if k in self._ARRAYS:
                self.set_array(k, v)
            elif k in ("centralSeedConfirmationRange", "forwardSeedConfirmationRange"):
                rng = getattr(self, k)
                for rk, rv in v.items():
                    setattr(rng, rk, rv)
            else:
                if not any(k == f[0] for f in self._fields_):
                    raise AttributeError(f"unknown config field {k}")
                setattr(self, k, v)
        return self


class OrthogonalOptions(C.Structure):
    """b200seed_orthogonal_options: the OrthogonalTripletSeedingAlgorithm::Config members the grid Config lacks."""
    _fields_ = [
        ("zOutermostLayersMin", C.c_float),
        ("zOutermostLayersMax", C.c_float),
        ("deltaPhiMax", C.c_float),
    ]


class Seeds(C.Structure):
    _fields_ = [
        ("bottom", C.c_void_p),
        ("middle", C.c_void_p),
        ("top", C.c_void_p),
        ("quality", C.c_void_p),
        ("vertexZ", C.c_void_p),
        ("capacity", C.c_uint64),
        ("size", C.c_uint64),
    ]


class Info(C.Structure):
    _fields_ = [
        ("phiBins", C.c_int32),
        ("zBins", C.c_int32),
        ("rBins", C.c_int32),
        ("nGlobalBins", C.c_int32),
        ("minHelixDiameter2", C.c_float),
        ("highland", C.c_float),
        ("sigmapT2perRadius", C.c_float),
        ("multipleScattering2", C.c_float),
        ("deltaRMinBottom", C.c_float),
        ("deltaRMaxBottom", C.c_float),
        ("deltaRMinTop", C.c_float),
        ("deltaRMaxTop", C.c_float),
        ("smCount", C.c_int32),
        ("ccMajor", C.c_int32),
        ("ccMinor", C.c_int32),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("nSpacePoints", C.c_uint64),
        ("nInGrid", C.c_uint64),
        ("nMiddles", C.c_uint64),
        ("nBottomDoublets", C.c_uint64),
        ("nTopDoublets", C.c_uint64),
        ("nTripletTests", C.c_uint64),
        ("nCandidates", C.c_uint64),
        ("nSeeds", C.c_uint64),
        ("nTieMiddles", C.c_uint64),
        ("nKernelLaunches", C.c_uint64),
        ("nConfirmationRounds", C.c_uint64),
    ]

    def as_dict(self):
        return {f[0]: int(getattr(self, f[0])) for f in self._fields_}


class Doublets(C.Structure):
    _fields_ = [
        ("nMiddles", C.c_uint64),
        ("nDoublets", C.c_uint64),
        ("middlePos", C.c_void_p),
        ("firstDoublet", C.c_void_p),
        ("nBottom", C.c_void_p),
        ("otherPos", C.c_void_p),
        ("cotTheta", C.c_void_p),
        ("iDeltaR", C.c_void_p),
        ("er", C.c_void_p),
        ("u", C.c_void_p),
        ("v", C.c_void_p),
        ("xNew", C.c_void_p),
        ("yNew", C.c_void_p),
        ("middleCapacity", C.c_uint64),
        ("doubletCapacity", C.c_uint64),
        ("gpuMilliseconds", C.c_float),
    ]


def f32(v: float) -> float:
    """Round a python float to binary32 (what assigning to a C float does)."""
    return C.c_float(v).value


def seeding_py_config(init) -> Config:
    """Config 1 of BASELINE.json: Examples/Scripts/Python/seeding.py:117-134 verbatim.

    ``init`` is a ``*_config_init(Config*)`` function (product or oracle).
    """
    cfg = Config()
    init(C.byref(cfg))
    cfg.update(
        rMax=200 * mm,
        deltaRMin=1 * mm,
        deltaRMax=300 * mm,
        deltaRMinTop=1 * mm,
        deltaRMaxTop=300 * mm,
        deltaRMinBottom=1 * mm,
        deltaRMaxBottom=300 * mm,
        collisionRegionMin=-250 * mm,
        collisionRegionMax=250 * mm,
        zMin=-2000 * mm,
        zMax=2000 * mm,
        maxSeedsPerSpM=1,
        sigmaScattering=50,
        radLengthPerSeed=0.1,
        minPt=500 * MeV,
        impactMax=3 * mm,
        bFieldInZ=2 * T,
    )
    return cfg


def pu200_config(init) -> Config:
    """Configs 2-5 of BASELINE.json: the reference's own <mu>=200 cut set,
    CI/physmon/workflows/physmon_trackfinding_ttbar_pu200.py:104-114
    (identical to seeding.py except sigmaScattering=5; rMin=33 is ignored by the
    grid, GridTripletSeedingAlgorithm.cpp:133-134)."""
    cfg = seeding_py_config(init)
    cfg.update(sigmaScattering=5, rMin=33 * mm)
    return cfg


def itk_like_config(init) -> Config:
    """A non-uniform z-binned configuration exercising zBinEdges,
    zBinNeighbors{Top,Bottom}, zBinsCustomLooping and rRangeMiddleSP
    (shape of Python/Examples/python/itk.py:300-560 scaled to the generic
    detector; used by parity tests only)."""
    cfg = pu200_config(init)
    edges = [-2000.0, -1000.0, -500.0, -200.0, 0.0, 200.0, 500.0, 1000.0, 2000.0]
    n = len(edges) - 1
    cfg.update(
        zBinEdges=edges,
        zBinNeighborsTop=[(0, 0), (-1, 0), (-1, 0), (-1, 0), (0, 1), (0, 1), (0, 1), (0, 0)],
        zBinNeighborsBottom=[(0, 1), (0, 1), (0, 1), (0, 1), (-1, 0), (-1, 0), (-1, 0), (-1, 0)],
        zBinsCustomLooping=[1, 8, 2, 7, 3, 6, 4, 5],
        rRangeMiddleSP=[(60.0, 130.0)] * 2 + [(50.0, 120.0)] * (n - 4) + [(60.0, 130.0)] * 2,
        numPhiNeighbors=1,
        maxSeedsPerSpM=4,
        deltaRMinTop=6.0,
        deltaRMaxTop=280.0,
        deltaRMinBottom=6.0,
        deltaRMaxBottom=150.0,
    )
    return cfg


def confirmation_overrides() -> dict:
    """The seed-confirmation block of the ITk pixel configuration
    (Python/Examples/python/itk.py:354-376,442-447)."""
    rng = dict(rMaxSeedConf=140.0, nTopForLargeR=1, nTopForSmallR=2, seedConfMinBottomRadius=60.0,
               seedConfMaxZOrigin=150.0, minImpactSeedConf=1.0)
    return dict(
        seedConfirmation=1,
        centralSeedConfirmationRange=dict(zMinSeedConf=-500.0, zMaxSeedConf=500.0, **rng),
        forwardSeedConfirmationRange=dict(zMinSeedConf=-3000.0, zMaxSeedConf=3000.0, **rng),
        zOriginWeightFactor=1.0,
        compatSeedWeight=100.0,
        impactWeightFactor=100.0,
        compatSeedLimit=3,
        numSeedIncrement=100.0,
        seedWeightIncrement=0.0,
        maxSeedsPerSpMConf=5,
        maxQualitySeedsPerSpMConf=5,
    )


def itk_conf_config(init) -> Config:
    """itk_like_config with seedConfirmation = true (the second published
    configuration's filter, itk.py:300-560, on the generic-detector geometry)."""
    return itk_like_config(init).update(**confirmation_overrides())


def itk_pixel_config(init, z_neighbors: bool = True, high_occupancy: bool = False) -> Config:
    """The reference's second published configuration VERBATIM: ``itkSeedingAlgConfig(PixelSpacePoints)``
    (Python/Examples/python/itk.py:302-560) as ``addGridTripletSeeding`` hands it to
    ``GridTripletSeedingAlgorithm::Config`` (Python/Examples/python/reconstruction.py:1001-1077): 13 non-uniform z
    bins out to +-3 m, the rRangeMiddleSP table, zBinsCustomLooping, interactionPointCut, seedConfirmation with
    both range blocks, useDeltaRinsteadOfTopRadius, deltaZMax = inf.

    ``z_neighbors``: also set zBinNeighborsTop / zBinNeighborsBottom / numPhiNeighbors (itk.py:410-441,380);
    ``addGridTripletSeeding`` itself does not forward them (the fields keep their defaults), the legacy
    ``addStandardSeeding`` path does (reconstruction.py:966-968) -- both shapes are tested.
    ``high_occupancy``: the highOccupancyConfig branch (itk.py:451-456,492-511) incl. useExtraCuts."""
    cfg = Config()
    init(C.byref(cfg))
    r_range = [(40.0, 90.0), (40.0, 90.0), (40.0, 200.0), (46.0, 200.0), (46.0, 200.0), (46.0, 250.0), (46.0, 250.0),
               (46.0, 250.0), (46.0, 200.0), (46.0, 200.0), (40.0, 200.0), (40.0, 90.0), (40.0, 90.0)]
    looping = [2, 3, 4, 5, 12, 11, 10, 9, 7, 6, 8]
    r_max, delta_r_max = 320.0, 280.0
    min_pt, col_min, col_max, variable = 900 * MeV, -200.0, 200.0, 1
    if high_occupancy:
        r_max, delta_r_max = 250.0, 200.0
        looping = [2, 10, 3, 9, 6, 4, 8, 5, 7]
        min_pt, col_min, col_max, variable = 1000 * MeV, -150.0, 150.0, 0
        r_range = [(40.0, 80.0), (40.0, 80.0), (40.0, 200.0), (70.0, 200.0), (70.0, 200.0), (70.0, 250.0), (70.0, 250.0),
                   (70.0, 250.0), (70.0, 200.0), (70.0, 200.0), (40.0, 200.0), (40.0, 80.0), (40.0, 80.0)]
    cfg.update(
        bFieldInZ=2 * T, minPt=min_pt, cotThetaMax=27.2899, impactMax=2.0,
        deltaRMin=20.0, deltaRMax=delta_r_max, deltaRMinTop=6.0, deltaRMaxTop=280.0, deltaRMinBottom=6.0, deltaRMaxBottom=150.0,
        rMax=r_max, zMin=-3000.0, zMax=3000.0, phiMin=-math.pi, phiMax=math.pi,
        phiBinDeflectionCoverage=3, maxPhiBins=200,
        zBinEdges=[-3000.0, -2700.0, -2500.0, -1400.0, -925.0, -500.0, -250.0, 250.0, 500.0, 925.0, 1400.0, 2500.0, 2700.0, 3000.0],
        zBinsCustomLooping=looping,
        useVariableMiddleSPRange=variable, rRangeMiddleSP=r_range,
        deltaRMiddleMinSPRange=10.0, deltaRMiddleMaxSPRange=10.0,
        interactionPointCut=1, collisionRegionMin=col_min, collisionRegionMax=col_max,
        sigmaScattering=2.0, radLengthPerSeed=0.0975,
        compatSeedWeight=100.0, impactWeightFactor=100.0, zOriginWeightFactor=1.0,
        maxSeedsPerSpM=4, compatSeedLimit=3, seedWeightIncrement=0.0, numSeedIncrement=100.0,
        seedConfirmation=1,
        centralSeedConfirmationRange=dict(zMinSeedConf=-500.0, zMaxSeedConf=500.0, rMaxSeedConf=140.0, nTopForLargeR=1,
                                          nTopForSmallR=2, seedConfMinBottomRadius=60.0, seedConfMaxZOrigin=150.0,
                                          minImpactSeedConf=1.0),
        forwardSeedConfirmationRange=dict(zMinSeedConf=-3000.0, zMaxSeedConf=3000.0, rMaxSeedConf=140.0, nTopForLargeR=1,
                                          nTopForSmallR=2, seedConfMinBottomRadius=60.0, seedConfMaxZOrigin=150.0,
                                          minImpactSeedConf=1.0),
        maxSeedsPerSpMConf=5, maxQualitySeedsPerSpMConf=5, useDeltaRinsteadOfTopRadius=1,
        useExtraCuts=1 if high_occupancy else 0,
    )
    if z_neighbors:
        cfg.update(
            zBinNeighborsTop=[(0, 0), (-1, 0), (-2, 0), (-1, 0), (-1, 0), (-1, 0), (-1, 1), (0, 1), (0, 1), (0, 1), (0, 2), (0, 1), (0, 0)],
            zBinNeighborsBottom=[(0, 0), (0, 1), (0, 1), (0, 1), (0, 1), (0, 1), (0, 0), (-1, 0), (-1, 0), (-1, 0), (-1, 0), (-1, 0), (0, 0)],
            numPhiNeighbors=1,
        )
    return cfg


def itk_strip_config(init, z_neighbors: bool = True) -> Config:
    """``itkSeedingAlgConfig(StripSpacePoints)`` VERBATIM (Python/Examples/python/itk.py:302-400,458-506) as
    ``addGridTripletSeeding`` hands it to ``GridTripletSeedingAlgorithm::Config`` (reconstruction.py:1001-1077):
    rMax 1200 mm, deltaR 20 - 600 mm with 20 - 300 mm per side, deltaZMax 900 mm, no interaction-point cut,
    impactMax 20 mm, the strip z-bin looping and neighbour tables, variable middle range 30 / 150 mm,
    seedConfirmation with collector capacities 100 / 100, seedWeightIncrement 10100, compatSeedLimit 4."""
    cfg = Config()
    init(C.byref(cfg))
    r_range = [(40.0, 90.0), (40.0, 90.0), (40.0, 200.0), (46.0, 200.0), (46.0, 200.0), (46.0, 250.0), (46.0, 250.0),
               (46.0, 250.0), (46.0, 200.0), (46.0, 200.0), (40.0, 200.0), (40.0, 90.0), (40.0, 90.0)]
    conf = dict(rMaxSeedConf=140.0, nTopForLargeR=1, nTopForSmallR=2, seedConfMinBottomRadius=60.0,
                seedConfMaxZOrigin=150.0, minImpactSeedConf=1.0)
    cfg.update(
        bFieldInZ=2 * T, minPt=900 * MeV, cotThetaMax=27.2899, impactMax=20.0,
        deltaRMin=20.0, deltaRMax=600.0, deltaRMinTop=20.0, deltaRMaxTop=300.0, deltaRMinBottom=20.0, deltaRMaxBottom=300.0,
        deltaZMax=900.0,
        rMax=1200.0, zMin=-3000.0, zMax=3000.0, phiMin=-math.pi, phiMax=math.pi,
        phiBinDeflectionCoverage=3, maxPhiBins=200,
        zBinEdges=[-3000.0, -2700.0, -2500.0, -1400.0, -925.0, -500.0, -250.0, 250.0, 500.0, 925.0, 1400.0, 2500.0, 2700.0, 3000.0],
        zBinsCustomLooping=[6, 7, 5, 8, 4, 9, 3, 10, 2, 11, 1],
        useVariableMiddleSPRange=1, rRangeMiddleSP=r_range,
        deltaRMiddleMinSPRange=30.0, deltaRMiddleMaxSPRange=150.0,
        interactionPointCut=0, collisionRegionMin=-200.0, collisionRegionMax=200.0,
        sigmaScattering=2.0, radLengthPerSeed=0.0975,
        compatSeedWeight=100.0, impactWeightFactor=1.0, zOriginWeightFactor=1.0,
        maxSeedsPerSpM=4, compatSeedLimit=4, seedWeightIncrement=10100.0, numSeedIncrement=1.0,
        seedConfirmation=1,
        centralSeedConfirmationRange=dict(zMinSeedConf=-500.0, zMaxSeedConf=500.0, **conf),
        forwardSeedConfirmationRange=dict(zMinSeedConf=-3000.0, zMaxSeedConf=3000.0, **conf),
        maxSeedsPerSpMConf=100, maxQualitySeedsPerSpMConf=100, useDeltaRinsteadOfTopRadius=0,
        useExtraCuts=0,
    )
    if z_neighbors:
        cfg.update(
            zBinNeighborsTop=[(0, 0), (-1, 0), (-2, 0), (-1, 0), (-1, 0), (-1, 0), (-1, 1), (0, 1), (0, 1), (0, 1), (0, 2), (0, 1), (0, 0)],
            zBinNeighborsBottom=[(0, 0), (0, 1), (0, 1), (0, 1), (0, 2), (0, 1), (0, 0), (-1, 0), (-2, 0), (-1, 0), (-1, 0), (-1, 0), (0, 0)],
            numPhiNeighbors=1,
        )
    return cfg


def orthogonal_config(init_orth, **overrides):
    """(Config, OrthogonalOptions) with the reference defaults of OrthogonalTripletSeedingAlgorithm::Config
    (hpp:38-186) and the <mu>=200 cut set of pu200_config on top (the same physics cuts through the other
    candidate provider; addOrthogonalTripletSeeding forwards the same SeedFinderConfigArg fields,
    Python/Examples/python/reconstruction.py:1080-1160).  ``init_orth`` is ``*_orthogonal_config_init``."""
    cfg, opt = Config(), OrthogonalOptions()
    init_orth(C.byref(cfg), C.byref(opt))
    cfg.update(
        rMax=200 * mm, deltaRMin=1 * mm, deltaRMax=300 * mm, deltaRMinTop=1 * mm, deltaRMaxTop=300 * mm,
        deltaRMinBottom=1 * mm, deltaRMaxBottom=300 * mm, collisionRegionMin=-250 * mm, collisionRegionMax=250 * mm,
        zMin=-2000 * mm, zMax=2000 * mm, maxSeedsPerSpM=1, sigmaScattering=5, radLengthPerSeed=0.1, minPt=500 * MeV,
        impactMax=3 * mm, bFieldInZ=2 * T,
    )
    opt_over = {k: overrides.pop(k) for k in list(overrides) if k in ("zOutermostLayersMin", "zOutermostLayersMax", "deltaPhiMax")}
    cfg.update(**overrides)
    for k, v in opt_over.items():
        setattr(opt, k, v)
    return cfg, opt


NAN = math.nan
