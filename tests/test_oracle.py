"""CPU tests of the oracle (the restatement of the reference algorithm).

The reference has no golden vector for this path, so the
oracle is anchored on (i) the worked constants of the canonical configuration,
(ii) the reference's own utility-layer known answers (GridBinFinderTests,
BinnedGroupTests, SpacePointGridPhiBinningTests), (iii) hand-checkable micro
cases, (iv) self-consistency properties and (v) the committed fixtures.
"""
import glob
import math
import os

import numpy as np
import pytest

from tests.conftest import make_config


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle

    return oracle


def test_worked_constants_of_canonical_config(O):
    """SURVEY.md section 3.4 (computed there with g++ float from the reference's expressions)."""
    i = O.Oracle(make_config("seeding_py", O.config_init)).info()
    assert i.phiBins == 53 and i.zBins == 1 and i.rBins == 1 and i.nGlobalBins == 55 * 3 * 3
    assert i.minHelixDiameter2 == np.float32(2781625.5)
    assert i.highland == np.float32(3.92439403e-3)
    assert i.sigmapT2perRadius == np.float32(428394.469)
    assert i.multipleScattering2 == np.float32(0.154008687)
    j = O.Oracle(make_config("pu200", O.config_init)).info()
    assert j.phiBins == 53
    assert abs(j.sigmapT2perRadius - 4283.94) < 0.01
    assert j.multipleScattering2 == np.float32(1.54008687e-3) or abs(j.multipleScattering2 - 1.54008687e-3) < 1e-9


def test_algorithm_default_config_has_26_phi_bins(O):
    from acts_b200.config import Config
    import ctypes as C

    cfg = Config()
    O.config_init(C.byref(cfg))
    assert O.Oracle(cfg).info().phiBins == 26


def test_phi_binning_reference_known_answers(O):
    """Tests/UnitTests/Core/Seeding/SpacePointGridPhiBinningTests.cpp:44-86."""
    from acts_b200 import config as cm

    cfg = make_config("seeding_py", O.config_init).update(bFieldInZ=0.0, maxPhiBins=123)
    assert O.Oracle(cfg).info().phiBins == 123          # zero field -> maxPhiBins
    cfg = make_config("seeding_py", O.config_init).update(maxPhiBins=7)
    assert O.Oracle(cfg).info().phiBins == 7             # cap
    cfg = make_config("seeding_py", O.config_init).update(minPt=0.010)
    with pytest.raises(O.OracleError) as ei:             # minPt = 10 MeV at rMax = 200 -> domain_error
        O.Oracle(cfg)
    assert ei.value.code == cm.ERR_DOMAIN


def test_axis_neighbourhood_reference_known_answers(O):
    """Tests/UnitTests/Core/Utilities/GridBinFinderTests.cpp:25-82, 10-bin axes, +-1."""
    def closed(idx):
        buf = np.zeros(64, np.uint64)
        n = O.lib().oracle_neighbors_closed(idx, -1, 1, 10, buf.ctypes.data, 64)
        return buf[:n].tolist()

    def opened(idx):
        buf = np.zeros(64, np.uint64)
        n = O.lib().oracle_neighbors_open(idx, -1, 1, 10, buf.ctypes.data, 64)
        return buf[:n].tolist()

    assert closed(1) == [10, 1, 2]
    assert closed(10) == [9, 10, 1]
    assert closed(5) == [4, 5, 6]
    assert opened(1) == [0, 1, 2]
    assert opened(10) == [9, 10, 11]


def test_navigation_validation(O):
    """BinnedGroupTests.cpp:72-120: custom visit order, rejection of 0 / nBins+1 / duplicates."""
    from acts_b200 import config as cm

    edges = [-2000.0, -500.0, 0.0, 500.0, 2000.0]
    ok = make_config("pu200", O.config_init).update(zBinEdges=edges, zBinsCustomLooping=[3, 4, 2, 1])
    O.Oracle(ok)
    for bad in ([0, 1, 2], [1, 2, 5 - 0], [1, 1, 2]):
        cfg = make_config("pu200", O.config_init).update(zBinEdges=edges, zBinsCustomLooping=bad)
        with pytest.raises(O.OracleError) as ei:
            O.Oracle(cfg)
        assert ei.value.code == cm.ERR_INVALID_ARGUMENT


def _helix_points(radii, pt, phi0, z0, cot, q=1.0):
    R = pt / (2 * 0.000299792458)
    xs, ys, zs = [], [], []
    for r in radii:
        alpha = 2 * math.asin(r / (2 * R))
        phi = phi0 - q * 0.5 * alpha
        xs.append(r * math.cos(phi)); ys.append(r * math.sin(phi)); zs.append(z0 + R * alpha * cot)
    x, y, z = (np.array(v, dtype=np.float32) for v in (xs, ys, zs))
    return {"x": x, "y": y, "z": z, "r": np.hypot(x.astype(np.float64), y.astype(np.float64)).astype(np.float32),
            "varZ": np.full(len(radii), 2e-4, np.float32), "varR": np.full(len(radii), 4e-6, np.float32)}


def test_three_points_on_a_helix_give_exactly_one_seed(O):
    """Hand-checkable micro case: a 2 GeV helix through the origin crossing r = 32, 72, 116:
    one seed (bottom, middle, top) = (0, 1, 2), quality = -impact (no compatible
    second top), vertexZ = zM - rM * cotTheta_bottom ~ z0."""
    ev = _helix_points([32.0, 72.0, 116.0], pt=2.0, phi0=0.3, z0=12.0, cot=0.5)
    res = O.Oracle(make_config("pu200", O.config_init)).run(ev)
    assert res["quality"].size == 1
    assert (int(res["bottom"][0]), int(res["middle"][0]), int(res["top"][0])) == (0, 1, 2)
    assert -0.05 < float(res["quality"][0]) <= 0.0          # -impact, impact ~ 0 for a track from the origin
    assert abs(float(res["vertexZ"][0]) - 12.0) < 0.5
    # a fourth point on the same helix is a compatible top: the weights gain compatSeedWeight
    ev4 = _helix_points([32.0, 72.0, 116.0, 172.0], pt=2.0, phi0=0.3, z0=12.0, cot=0.5)
    res4 = O.Oracle(make_config("pu200", O.config_init)).run(ev4)
    triplets = {(int(b), int(m), int(t)): float(q) for b, m, t, q in zip(res4["bottom"], res4["middle"], res4["top"], res4["quality"])}
    assert (0, 1, 2) in triplets and (0, 1, 3) in triplets
    assert 199.9 < triplets[(0, 1, 2)] <= 200.0 and 199.9 < triplets[(0, 1, 3)] <= 200.0


def test_low_pt_helix_is_rejected(O):
    ev = _helix_points([32.0, 72.0, 116.0], pt=0.2, phi0=-1.0, z0=0.0, cot=0.2)  # below minPt = 0.5
    assert O.Oracle(make_config("pu200", O.config_init)).run(ev)["quality"].size == 0


def test_seeds_satisfy_the_cuts_in_double_precision(O):
    """Self-consistency: every emitted seed passes the r-window, collision-region,
    cotTheta and helix/impact cuts when re-checked in float64 (with slack)."""
    from acts_b200 import events

    cfg = make_config("pu200", O.config_init)
    ev = events.pileup_event(4, mu=20)
    res = O.Oracle(cfg).run(ev)
    assert res["quality"].size > 1000
    x, y, z, r = (ev[k].astype(np.float64) for k in ("x", "y", "z", "r"))
    b, m, t = res["bottom"], res["middle"], res["top"]
    assert np.all((r[m] >= 60 - 1e-3) & (r[m] <= 120 + 1e-3))
    for o, sign in ((b, -1.0), (t, 1.0)):
        dR = sign * (r[o] - r[m])
        dZ = sign * (z[o] - z[m])
        assert np.all((dR >= 1 - 1e-3) & (dR <= 300 + 1e-3))
        z0 = z[m] - r[m] * dZ / dR
        assert np.all(np.abs(z0) <= 250 * (1 + 1e-4))
        assert np.all(np.abs(dZ / dR) <= 10.01788 * (1 + 1e-4))
    # circle through the three points: radius above minPt / (0.3 B), impact below impactMax
    ax, ay, bx, by, cx, cy = x[b], y[b], x[m], y[m], x[t], y[t]
    d = 2 * (ax * (by - cy) + bx * (cy - ay) + cx * (ay - by))
    ux = ((ax**2 + ay**2) * (by - cy) + (bx**2 + by**2) * (cy - ay) + (cx**2 + cy**2) * (ay - by)) / d
    uy = ((ax**2 + ay**2) * (cx - bx) + (bx**2 + by**2) * (ax - cx) + (cx**2 + cy**2) * (bx - ax)) / d
    rad = np.hypot(ax - ux, ay - uy)
    assert np.all(rad >= 0.5 / (2 * 0.000299792458) * 0.98)
    assert np.all(np.abs(np.hypot(ux, uy) - rad) <= 3.0 * 1.05)
    # per middle at most maxSeedsPerSpM + 1 seeds
    _, counts = np.unique(m, return_counts=True)
    assert counts.max() <= cfg.maxSeedsPerSpM + 1


def test_stable_and_faithful_sort_orders_give_the_same_seed_set(O):
    from acts_b200 import events

    orc = O.Oracle(make_config("pu200", O.config_init))
    ev = events.pileup_event(1, mu=40)
    assert O.seed_set(orc.run(ev, sort_mode=O.Oracle.FAITHFUL)) == O.seed_set(orc.run(ev, sort_mode=O.Oracle.STABLE))


def test_oracle_reproduces_committed_golden_vectors(O):
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
    assert len(files) >= 5
    for f in files:
        g = np.load(f)
        ev = {k: g[k] for k in ("x", "y", "z", "r", "varZ", "varR")}
        res = O.Oracle(make_config(str(g["config"]), O.config_init)).run(ev, want_grid=True)
        for k in ("bottom", "middle", "top", "quality", "vertexZ"):
            assert np.array_equal(res[k].view(np.uint32), g[k].view(np.uint32)), (f, k)
        assert np.array_equal(res["grid"]["copiedFromIndex"], g["grid_copiedFromIndex"])
        if "strip" in g.files:  # the strip triplet path on the same event
            ev["strip"] = g["strip"]
            res = O.Oracle(make_config(str(g["config"]), O.config_init)).run(ev, strip_cot_theta_diff_max=float(g["s_cotThetaDiffMax"]))
            for k in ("bottom", "middle", "top", "quality", "vertexZ"):
                assert np.array_equal(res[k].view(np.uint32), g["s_" + k].view(np.uint32)), (f, "strip", k)


def test_empty_and_out_of_grid_inputs(O):
    orc = O.Oracle(make_config("pu200", O.config_init))
    empty = {k: np.zeros(0, np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")}
    assert orc.run(empty)["quality"].size == 0
    ev = _helix_points([32.0, 72.0, 116.0], pt=2.0, phi0=0.3, z0=12.0, cot=0.5)
    ev["r"][2] = 250.0  # outside rMax = 200: dropped, no seed possible
    assert orc.run(ev)["quality"].size == 0


def test_estimated_parameters_recover_the_generated_helix(O):
    """f1 oracle (Acts::estimateTrackParamsFromSeed): a seed on an ideal helix gives back its
    direction and charge-signed curvature."""
    pt, phi0, z0, cot = 2.0, 0.3, 12.0, 0.5
    for q in (1.0, -1.0):
        ev = _helix_points([32.0, 72.0, 116.0], pt=pt, phi0=phi0, z0=z0, cot=cot, q=q)
        seeds = {"bottom": np.array([0], np.uint32), "middle": np.array([1], np.uint32), "top": np.array([2], np.uint32)}
        par = O.estimate_params(seeds, ev)[0]
        assert np.allclose(par[:3], [ev["x"][0], ev["y"][0], ev["z"][0]])
        assert abs(np.linalg.norm(par[4:7]) - 1.0) < 1e-12
        p = pt * math.sqrt(1 + cot * cot)
        assert abs(abs(par[7]) - 1.0 / p) < 2e-3 / p           # |q/p|
        assert abs(par[6] / math.hypot(par[4], par[5]) - cot) < 2e-3   # dz/ds = cot(theta)
        # the two charges bend in opposite directions
        if q > 0:
            qp_plus = par[7]
        else:
            assert qp_plus * par[7] < 0


def test_pixel_space_point_maker_against_independent_float64(O):
    """f4: createPixelSpacePoint (SpacePointMaker.cpp:44-76) restated in the oracle vs a numpy float64 evaluation
    of the same formulas (global = T (l0, l1, 0, 1); J = d(z, r)/d(x, y, z) R[:, :2]; diag(J C J^T))."""
    from acts_b200 import events

    meas, tr = events.pixel_measurements(3, n=5000)
    got = O.make_pixel_spacepoints(meas, tr)
    T = tr[meas["surface"]]
    loc = np.stack([meas["loc0"], meas["loc1"], np.zeros_like(meas["loc0"]), np.ones_like(meas["loc0"])], axis=1)
    g = np.einsum("nij,nj->ni", T, loc)
    rr = np.hypot(g[:, 0], g[:, 1])
    J = np.zeros((g.shape[0], 2, 3))
    J[:, 0, 2] = 1
    J[:, 1, 0] = g[:, 0] / rr
    J[:, 1, 1] = g[:, 1] / rr
    jac = np.einsum("nak,nkb->nab", J, T[:, :, :2])
    Cm = np.stack([np.stack([meas["cov00"], meas["cov01"]], 1), np.stack([meas["cov01"], meas["cov11"]], 1)], 1)
    cov = np.einsum("nab,nbc,ndc->nad", jac, Cm, jac)
    for k, ref in (("x", g[:, 0]), ("y", g[:, 1]), ("z", g[:, 2]), ("r", rr), ("varZ", cov[:, 0, 0]), ("varR", cov[:, 1, 1])):
        r32 = ref.astype(np.float32)
        assert np.all(np.abs(got[k] - r32) <= np.spacing(np.abs(r32))), k   # float32 columns: at most one ulp apart
        assert np.mean(got[k] == r32) > 0.999, k
    # barrel modules: varZ is the local-y variance, varR tiny (tilt only); a hand-checkable case
    one = {"surface": np.array([0], np.uint32), "loc0": np.array([0.0]), "loc1": np.array([0.0]),
           "cov00": np.array([4.0]), "cov01": np.array([0.0]), "cov11": np.array([9.0])}
    sp = O.make_pixel_spacepoints(one, tr)
    assert sp["varZ"][0] == 9.0 and abs(sp["varR"][0] - 4.0 * np.sin(0.14) ** 2) < 1e-6
    assert abs(sp["r"][0] - 32.0) < 1e-5 and sp["z"][0] == -468.0


def test_estimate_params_reference_known_answer(O):
    """Tests/UnitTests/Core/Seeding/EstimateTrackParamsFromSeedTest.cpp:187-195 (trackparm_estimate_aligined):
    three aligned space points give q/p == 0 exactly."""
    ev = {"x": np.array([-72.775, -84.325, -98.175], np.float32), "y": np.array([-0.325, -0.325, -0.325], np.float32),
          "z": np.array([-615.6, -715.6, -835.6], np.float32)}
    seeds = {"bottom": np.array([0], np.uint32), "middle": np.array([1], np.uint32), "top": np.array([2], np.uint32)}
    out = O.estimate_params(seeds, ev, b_field=(0.0, 0.0, 0.000899377))
    assert out[0, 7] == 0.0
    assert not np.isnan(out).any()


def test_synthetic_strip_details_calibrate_back_to_the_space_point():
    """acts_b200.events.strip_details models a double-sided module: running the reference's calibration formula
    (StripSpacePointCalibrationImpl.hpp:44-74: scale = d . (ihv x ohv), sOuter = d . (iosv x ihv),
    result = oc + ohv * sOuter / scale) with the direction from the origin gives the space point back, inside both
    strips (|s| <= 1) -- i.e. the synthetic column is a consistent geometry, not noise."""
    from acts_b200 import events

    for ev in (events.pileup_event(3, mu=20), events.itk_strip_event(1, mu=20)):
        d = (ev["strip"] if "strip" in ev else events.strip_details(ev, seed=3)).astype(np.float64)
        oc, iosv, ohv, ihv = d[:, 0:3], d[:, 3:6], d[:, 6:9], d[:, 9:12]
        p = np.stack([ev["x"], ev["y"], ev["z"]], axis=1).astype(np.float64)
        direction = p / np.linalg.norm(p, axis=1, keepdims=True)
        scale = np.sum(direction * np.cross(ihv, ohv), axis=1)
        s_inner = np.sum(direction * np.cross(iosv, ohv), axis=1) / scale
        s_outer = np.sum(direction * np.cross(iosv, ihv), axis=1) / scale
        assert np.all(np.abs(s_inner) <= 0.95) and np.all(np.abs(s_outer) <= 0.95)
        back = oc + ohv * s_outer[:, None]
        assert np.max(np.abs(back - p)) < 2e-3  # mm (float32 details)
