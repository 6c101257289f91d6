"""Synthetic Generic-detector-like space point events (numpy, deterministic).

There is no Fatras in this image, so the benchmark events are generated from
ideal helices through a pixel layout that follows the reference's Generic
detector (Examples/Detectors/GenericDetector/src/GenericDetectorBuilder.cpp
:328-335,364-368,405-406): barrel cylinders at r = 32/72/116/172 mm with
|z| <= 490 mm and a +-1 mm radial stagger (tilted modules), end-cap discs at
|z| = 600/700/820/960 mm with r in [30, 176] mm.  2 T solenoid field, pile-up
vertices z ~ N(0, 55.5 mm) like CI/physmon/workflows/
physmon_trackfinding_ttbar_pu200.py:60-63, Gaussian smearing sigma = 0.0144338 mm
(Examples/Configs/generic-digi-smearing-config.json).  Columns are binary32
exactly like the reference's SpacePointContainer (SpacePointMaker.cpp:260-264).

RNG: Philox counter-based, key = 42 + event number (the reference's seed
convention, Examples/Scripts/Python/seeding.py:69).
"""
from __future__ import annotations

import numpy as np

BARREL_R = np.array([32.0, 72.0, 116.0, 172.0])
BARREL_HALF_Z = 490.0
ENDCAP_Z = np.array([600.0, 700.0, 820.0, 960.0])
ENDCAP_R = (30.0, 176.0)
SIGMA = 0.0144338
B_FIELD_T = 2.0
PT_PER_RADIUS = B_FIELD_T * 0.000299792458  # GeV / mm


GENERIC_LAYOUT = dict(barrel_r=BARREL_R, barrel_half_z=BARREL_HALF_Z, endcap_z=ENDCAP_Z, endcap_r=ENDCAP_R)
# ITk-shaped pixel layout (five barrel layers, end-cap rings out to |z| = 2.85 m, r up to 315 mm): the shape the
# reference's second published configuration is written for (Python/Examples/python/itk.py:302-560: zBinEdges out to
# +-3000 mm, rMax 320 mm, collision region +-200 mm).  Not the real ITk geometry -- a layout with its extent, so that
# every z bin of the ITk table, the rRangeMiddleSP rows and the forward seedConfirmation range are populated.
ITK_LAYOUT = dict(barrel_r=np.array([34.0, 99.0, 160.0, 228.0, 291.0]), barrel_half_z=380.0,
                  endcap_z=np.array([450.0, 550.0, 680.0, 840.0, 1040.0, 1280.0, 1560.0, 1900.0, 2300.0, 2650.0, 2850.0]),
                  endcap_r=(33.0, 315.0))


# ITk-STRIP-shaped layout: four double-sided barrel layers out to 1 m, six end-cap discs out to |z| = 2.85 m -- the
# extent itk.py's StripSpacePoints block is written for (rMax 1000 / 1200 mm, deltaR up to 600 mm, deltaZMax 900 mm).
ITK_STRIP_LAYOUT = dict(barrel_r=np.array([405.0, 562.0, 762.0, 1000.0]), barrel_half_z=1372.0,
                        endcap_z=np.array([1512.0, 1702.0, 1952.0, 2252.0, 2602.0, 2852.0]),
                        endcap_r=(385.0, 970.0))


def _helix_hits(rng, zv, pt, eta, phi0, q, layout=None):
    """Intersections of helices from (0, 0, zv) with all layers (vectorised).

    Returns x, y, z (float64), is_barrel for every hit."""
    layout = GENERIC_LAYOUT if layout is None else layout
    BARREL_R, BARREL_HALF_Z = layout["barrel_r"], layout["barrel_half_z"]
    ENDCAP_Z, ENDCAP_R = layout["endcap_z"], layout["endcap_r"]
    R = pt / PT_PER_RADIUS  # mm
    sinh_eta = np.sinh(eta)
    xs, ys, zs, barrel = [], [], [], []
    # barrel: solve r = 2 R sin(alpha / 2)
    for r0 in BARREL_R:
        r = r0 + rng.uniform(-1.0, 1.0, size=pt.shape)
        ok = r < 2.0 * R * 0.999
        alpha = 2.0 * np.arcsin(np.where(ok, r / (2.0 * R), 0.0))
        z = zv + R * alpha * sinh_eta
        ok &= np.abs(z) <= BARREL_HALF_Z
        phi = phi0 - q * 0.5 * alpha
        # smear r-phi and z
        phi = phi + rng.normal(0.0, SIGMA, size=pt.shape) / r
        z = z + rng.normal(0.0, SIGMA, size=pt.shape)
        xs.append((r * np.cos(phi))[ok])
        ys.append((r * np.sin(phi))[ok])
        zs.append(z[ok])
        barrel.append(np.ones(int(ok.sum()), dtype=bool))
    # end-caps: solve z = z_disc
    for zd0 in ENDCAP_Z:
        for side in (-1.0, 1.0):
            zd = side * zd0 + rng.uniform(-2.0, 2.0, size=pt.shape)
            with np.errstate(divide="ignore", invalid="ignore"):
                s = (zd - zv) / sinh_eta
            alpha = s / R
            ok = np.isfinite(s) & (s > 0) & (alpha < np.pi)
            r = 2.0 * R * np.abs(np.sin(0.5 * np.where(ok, alpha, 0.0)))
            ok &= (r >= ENDCAP_R[0]) & (r <= ENDCAP_R[1])
            r = r + rng.normal(0.0, SIGMA, size=pt.shape)
            phi = phi0 - q * 0.5 * alpha + rng.normal(0.0, SIGMA, size=pt.shape) / np.maximum(r, 1.0)
            xs.append((r * np.cos(phi))[ok])
            ys.append((r * np.sin(phi))[ok])
            zs.append(zd[ok])
            barrel.append(np.zeros(int(ok.sum()), dtype=bool))
    return (np.concatenate(xs), np.concatenate(ys), np.concatenate(zs), np.concatenate(barrel))


def _finish(rng, x, y, z, barrel, noise_fraction, layout=None):
    layout = GENERIC_LAYOUT if layout is None else layout
    BARREL_R, BARREL_HALF_Z = layout["barrel_r"], layout["barrel_half_z"]
    ENDCAP_Z, ENDCAP_R = layout["endcap_z"], layout["endcap_r"]
    n = x.size
    n_noise = int(round(noise_fraction * n))
    if n_noise > 0:
        nb = n_noise // 2
        rb = rng.choice(BARREL_R, size=nb) + rng.uniform(-1.0, 1.0, size=nb)
        pb = rng.uniform(-np.pi, np.pi, size=nb)
        zb = rng.uniform(-BARREL_HALF_Z, BARREL_HALF_Z, size=nb)
        ne = n_noise - nb
        re = rng.uniform(ENDCAP_R[0], ENDCAP_R[1], size=ne)
        pe = rng.uniform(-np.pi, np.pi, size=ne)
        ze = rng.choice(ENDCAP_Z, size=ne) * rng.choice([-1.0, 1.0], size=ne) + rng.uniform(-2.0, 2.0, size=ne)
        x = np.concatenate([x, rb * np.cos(pb), re * np.cos(pe)])
        y = np.concatenate([y, rb * np.sin(pb), re * np.sin(pe)])
        z = np.concatenate([z, zb, ze])
        barrel = np.concatenate([barrel, np.ones(nb, dtype=bool), np.zeros(ne, dtype=bool)])
    perm = rng.permutation(x.size)
    x, y, z, barrel = x[perm], y[perm], z[perm], barrel[perm]
    r = np.hypot(x, y)
    var_hit = np.float32(SIGMA * SIGMA)
    varZ = np.where(barrel, var_hit, np.float32(0.0)).astype(np.float32)
    varR = np.where(barrel, np.float32(4e-6), var_hit).astype(np.float32)
    return {
        "x": x.astype(np.float32),
        "y": y.astype(np.float32),
        "z": z.astype(np.float32),
        "r": r.astype(np.float32),
        "varZ": varZ,
        "varR": varR,
    }


def pileup_event(event: int, mu: float = 200.0, sp_per_vertex: float = 500.0,
                 noise_fraction: float = 0.10, seed: int = 42) -> dict:
    """<mu> pile-up vertices of soft charged pions (configs 2-5).

    ``sp_per_vertex`` tunes the multiplicity so that <mu>=200 gives ~1e5 space
    points per event (BASELINE.json configs[2])."""
    rng = np.random.Generator(np.random.Philox(key=seed + event))
    n_vtx = max(1, int(rng.poisson(mu)))
    # ~4.5 hits per generated particle on this layout (measured)
    n_part = rng.poisson(sp_per_vertex / (1.0 + noise_fraction) / 4.5, size=n_vtx)
    zv = np.repeat(rng.normal(0.0, 55.5, size=n_vtx), n_part)
    n = zv.size
    pt = 0.1 + rng.gamma(2.0, 0.3, size=n)  # falling spectrum, GeV
    eta = rng.uniform(-2.7, 2.7, size=n)
    phi0 = rng.uniform(-np.pi, np.pi, size=n)
    q = rng.choice([-1.0, 1.0], size=n)
    x, y, z, barrel = _helix_hits(rng, zv, pt, eta, phi0, q)
    return _finish(rng, x, y, z, barrel, noise_fraction)


def itk_pileup_event(event: int, mu: float = 60.0, sp_per_vertex: float = 500.0,
                     noise_fraction: float = 0.10, seed: int = 42) -> dict:
    """Pile-up event on the ITk-shaped pixel layout (|eta| < 4, beam spot sigma_z = 50 mm)."""
    rng = np.random.Generator(np.random.Philox(key=seed + 100003 * 7 + event))
    n_vtx = max(1, int(rng.poisson(mu)))
    n_part = rng.poisson(sp_per_vertex / (1.0 + noise_fraction) / 5.5, size=n_vtx)
    zv = np.repeat(rng.normal(0.0, 50.0, size=n_vtx), n_part)
    n = zv.size
    pt = 0.1 + rng.gamma(2.0, 0.45, size=n)
    eta = rng.uniform(-4.0, 4.0, size=n)
    phi0 = rng.uniform(-np.pi, np.pi, size=n)
    q = rng.choice([-1.0, 1.0], size=n)
    x, y, z, barrel = _helix_hits(rng, zv, pt, eta, phi0, q, ITK_LAYOUT)
    return _finish(rng, x, y, z, barrel, noise_fraction, ITK_LAYOUT)


def itk_strip_event(event: int, mu: float = 60.0, sp_per_vertex: float = 300.0,
                    noise_fraction: float = 0.10, seed: int = 42) -> dict:
    """Pile-up event on the ITk-strip-shaped layout (|eta| < 2.7), with the strip calibration details of
    ``strip_details`` under ``"strip"`` (the points are what a strip space-point builder would hand over)."""
    rng = np.random.Generator(np.random.Philox(key=seed + 100003 * 11 + event))
    n_vtx = max(1, int(rng.poisson(mu)))
    n_part = rng.poisson(sp_per_vertex / (1.0 + noise_fraction) / 4.0, size=n_vtx)
    zv = np.repeat(rng.normal(0.0, 50.0, size=n_vtx), n_part)
    n = zv.size
    pt = 0.4 + rng.gamma(2.0, 0.6, size=n)
    eta = rng.uniform(-2.7, 2.7, size=n)
    phi0 = rng.uniform(-np.pi, np.pi, size=n)
    q = rng.choice([-1.0, 1.0], size=n)
    x, y, z, barrel = _helix_hits(rng, zv, pt, eta, phi0, q, ITK_STRIP_LAYOUT)
    ev = dict(_finish(rng, x, y, z, barrel, noise_fraction, ITK_STRIP_LAYOUT))
    ev["strip"] = strip_details(ev, seed=event, half_length=24.0, gap=6.0, barrel_max_z=1400.0)
    return ev


def muon_gun_event(event: int, n_muons: int = 100, seed: int = 42) -> dict:
    """Config 1: particle gun, 100 muons/event, pT 1-10 GeV, |eta| < 2.5
    (Examples/Scripts/Python/seeding.py:76-92), single vertex, no noise."""
    rng = np.random.Generator(np.random.Philox(key=seed + event))
    zv = np.full(n_muons, rng.normal(0.0, 55.5))
    pt = rng.uniform(1.0, 10.0, size=n_muons)
    eta = rng.uniform(-2.5, 2.5, size=n_muons)
    phi0 = rng.uniform(-np.pi, np.pi, size=n_muons)
    q = rng.choice([-1.0, 1.0], size=n_muons)
    x, y, z, barrel = _helix_hits(rng, zv, pt, eta, phi0, q)
    return _finish(rng, x, y, z, barrel, 0.0)


def concat_events(events: list[dict]) -> tuple[dict, np.ndarray]:
    """Concatenate events into the batch layout of ``b200seed_run_batch``."""
    offsets = np.zeros(len(events) + 1, dtype=np.uint32)
    for i, e in enumerate(events):
        offsets[i + 1] = offsets[i] + e["x"].size
    cols = {k: np.ascontiguousarray(np.concatenate([e[k] for e in events])) for k in ("x", "y", "z", "r", "varZ", "varR")}
    return cols, offsets


def pixel_measurements(event: int, n: int = 20000, seed: int = 42) -> tuple[dict, np.ndarray]:
    """Synthetic pixel measurements on generic-detector-like planar modules (input of the f4 space point maker).

    Barrel modules are planes tangent to the cylinders r = {32, 72, 116, 172} mm, tilted by 0.14 rad around the
    z axis (GenericDetectorBuilder.cpp:328-335); endcap modules are planes perpendicular to z at
    |z| = {600, 700, 820, 960} mm.  Local frame: x along the measurement (r-phi) direction, y along z (barrel) or r
    (endcap).  Returns (measurements, transforms): surface index, local position (float64), local covariance
    (pitch^2/12 on the diagonal, a small correlation so that every covariance term is exercised), and the
    (nSurfaces, 3, 4) local->global affine matrices.
    """
    rng = np.random.default_rng(seed + 7919 * event + 13)
    transforms = []
    for rl in (32.0, 72.0, 116.0, 172.0):
        n_phi = int(2 * np.pi * rl / 14.0)
        for k in range(n_phi):
            phi = 2 * np.pi * k / n_phi
            for zc in np.arange(-468.0, 469.0, 72.0):
                normal_phi = phi + 0.14
                ux = np.array([-np.sin(normal_phi), np.cos(normal_phi), 0.0])  # local x
                uy = np.array([0.0, 0.0, 1.0])                                  # local y
                uz = np.cross(ux, uy)
                centre = np.array([rl * np.cos(phi), rl * np.sin(phi), zc])
                transforms.append(np.column_stack([ux, uy, uz, centre]))
    for zd in (600.0, 700.0, 820.0, 960.0):
        for sign in (-1.0, 1.0):
            for k in range(40):
                phi = 2 * np.pi * k / 40
                for rc in (60.0, 130.0):
                    ux = np.array([-np.sin(phi), np.cos(phi), 0.0])
                    uy = np.array([np.cos(phi), np.sin(phi), 0.0])
                    uz = np.cross(ux, uy) * sign
                    centre = np.array([rc * np.cos(phi), rc * np.sin(phi), sign * zd])
                    transforms.append(np.column_stack([ux, sign * uy, uz, centre]))
    transforms = np.ascontiguousarray(np.stack(transforms), dtype=np.float64)
    ns = transforms.shape[0]
    surface = rng.integers(0, ns, n).astype(np.uint32)
    loc0 = rng.uniform(-8.4, 8.4, n)
    loc1 = rng.uniform(-36.0, 36.0, n)
    var = 0.0144338 ** 2 * rng.uniform(0.5, 4.0, (n, 2))
    rho = rng.uniform(-0.3, 0.3, n)
    meas = {"surface": surface, "loc0": loc0, "loc1": loc1, "cov00": var[:, 0], "cov11": var[:, 1],
            "cov01": rho * np.sqrt(var[:, 0] * var[:, 1])}
    return meas, transforms


def strip_details(ev: dict, seed: int = 0, half_length: float = 12.0, gap: float = 3.0, stereo: float = 0.02,
                  barrel_max_z: float = 500.0) -> np.ndarray:
    """Synthetic outer-strip calibration details for every space point of ``ev``: (n, 12) float32 =
    outerCenter, innerToOuterSeparation, outerHalfVector, innerHalfVector
    (Core/include/Acts/EventData/StripSpacePointCalibrationDetails.hpp:16-29).

    Every space point is taken as the crossing of a double-sided strip module: the outer strip passes through the
    point, the inner strip through the point where the line from the origin pierces a plane ``gap`` mm closer to the
    beam line; the two strips are rotated by +-``stereo`` rad about the module normal (radial in the barrel, along z
    for |z| > ``barrel_max_z``).  The crossing sits at a random place along both strips (|s| <= 0.9 half lengths), so a
    straight track from the origin calibrates back to the space point and tracks from displaced vertices move along
    the outer strip or leave the module (the tolerance cut of StripSpacePointCalibrationImpl.hpp:44-74)."""
    rng = np.random.Generator(np.random.Philox(key=977 + seed))
    p = np.stack([ev["x"], ev["y"], ev["z"]], axis=1).astype(np.float64)
    n = p.shape[0]
    r = np.hypot(p[:, 0], p[:, 1])
    barrel = np.abs(p[:, 2]) <= barrel_max_z
    rs = np.where(r > 0, r, 1.0)
    normal = np.where(barrel[:, None], np.stack([p[:, 0] / rs, p[:, 1] / rs, np.zeros(n)], axis=1),
                      np.tile(np.array([0.0, 0.0, 1.0]), (n, 1)))
    # strip axis without stereo: along z in the barrel, radial on the discs
    axis = np.where(barrel[:, None], np.tile(np.array([0.0, 0.0, 1.0]), (n, 1)),
                    np.stack([p[:, 0] / rs, p[:, 1] / rs, np.zeros(n)], axis=1))
    side = np.cross(normal, axis)

    def rotated(angle):
        return axis * np.cos(angle) + side * np.sin(angle)

    u_outer, u_inner = rotated(stereo), rotated(-stereo)
    s_outer = rng.uniform(-0.9, 0.9, size=n)[:, None]
    s_inner = rng.uniform(-0.9, 0.9, size=n)[:, None]
    pn = np.sum(p * normal, axis=1)
    pn = np.where(np.abs(pn) > 1e-6, pn, 1.0)
    q = p * (1.0 - gap / pn)[:, None]  # origin -> p pierces the inner plane here
    oc = p - s_outer * half_length * u_outer
    ic = q - s_inner * half_length * u_inner
    out = np.concatenate([oc, oc - ic, half_length * u_outer, half_length * u_inner], axis=1)
    return np.ascontiguousarray(out, dtype=np.float32)
