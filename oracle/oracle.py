"""ctypes wrapper of the CPU oracle (``oracle/seeding_oracle.cpp``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The
product package ``acts_b200`` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from acts_b200.config import Config, Info, OrthogonalOptions

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libseeding_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "seeding_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "acts_b200_seeding.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


class OracleCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "nSpacePoints", "nInGrid", "nMiddles", "nPairTests", "nBottomDoublets",
        "nTopDoublets", "nTripletTests", "nCandidates", "nSeeds", "nRTieBins",
        "nCotTieMiddles", "nCurvTieGroups", "nWeightTieMiddles", "maxBottoms",
        "maxTops", "maxCandidatesPerBottom", "maxCandidatesPerMiddle", "maxBinSize")]

    def as_dict(self):
        return {f[0]: int(getattr(self, f[0])) for f in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_config_init.argtypes = [C.POINTER(Config)]
        L.oracle_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.oracle_z_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.oracle_find_bins.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32]
        L.oracle_neighbors_closed.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
        L.oracle_neighbors_open.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
        L.oracle_bin_index.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.oracle_bin_index.restype = C.c_int64
        L.oracle_atan2f.argtypes = [C.c_float, C.c_float]
        L.oracle_atan2f.restype = C.c_float
        L.oracle_run.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.oracle_run_strips.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 7 + [C.c_float, C.c_int, C.POINTER(C.c_void_p)]
        L.oracle_result_free.argtypes = [C.c_void_p]
        L.oracle_result_num_seeds.argtypes = [C.c_void_p]
        L.oracle_result_num_seeds.restype = C.c_uint64
        L.oracle_result_seeds.argtypes = [C.c_void_p] * 6
        L.oracle_result_counters.argtypes = [C.c_void_p, C.POINTER(OracleCounters)]
        L.oracle_result_histograms.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_result_grid_size.argtypes = [C.c_void_p]
        L.oracle_result_grid_size.restype = C.c_uint64
        L.oracle_result_grid.argtypes = [C.c_void_p] * 10
        L.oracle_result_dump_middles.argtypes = [C.c_void_p]
        L.oracle_result_dump_middles.restype = C.c_uint64
        L.oracle_result_dump_doublets.argtypes = [C.c_void_p]
        L.oracle_result_dump_doublets.restype = C.c_uint64
        L.oracle_result_dump.argtypes = [C.c_void_p] * 12
        L.oracle_estimate_params.argtypes = [C.c_uint64] + [C.c_void_p] * 8
        L.oracle_make_pixel_spacepoints.argtypes = [C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32] + [C.c_void_p] * 7
        L.oracle_make_pixel_spacepoints.restype = C.c_int
        L.oracle_vertex_windows.argtypes = [C.POINTER(Config), C.c_uint32] + [C.c_void_p] * 4
        L.oracle_run_many.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 7 + [C.c_int, C.c_uint32, C.c_void_p]
        L.oracle_run_many.restype = C.c_int64
        L.oracle_orthogonal_config_init.argtypes = [C.POINTER(Config), C.POINTER(OrthogonalOptions)]
        L.oracle_create_orthogonal.argtypes = [C.POINTER(Config), C.POINTER(OrthogonalOptions), C.POINTER(C.c_void_p)]
        L.oracle_result_tree_order.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_result_tree_order.restype = C.c_uint64
        _lib = L
    return _lib


def config_init(cfg_ref):
    lib().oracle_config_init(cfg_ref)


def orthogonal_config_init(cfg_ref, opt_ref):
    lib().oracle_orthogonal_config_init(cfg_ref, opt_ref)


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """CPU restatement of GridTripletSeedingAlgorithm (reference semantics)."""

    FAITHFUL, STABLE = 0, 1

    def __init__(self, cfg: Config, orthogonal: OrthogonalOptions | None = None):
        """``orthogonal`` given: the restatement of OrthogonalTripletSeedingAlgorithm instead."""
        self._h = C.c_void_p()
        self._cfg = cfg
        self._orth = orthogonal
        if orthogonal is not None:
            rc = lib().oracle_create_orthogonal(C.byref(cfg), C.byref(orthogonal), C.byref(self._h))
        else:
            rc = lib().oracle_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise OracleError(rc, lib().oracle_last_error().decode())

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> Info:
        i = Info()
        lib().oracle_get_info(self._h, C.byref(i))
        return i

    def z_edges(self):
        buf = np.zeros(4096, dtype=np.float64)
        n = lib().oracle_z_edges(self._h, _p(buf), buf.size)
        return buf[:n].copy()

    def find_bins(self, phi_loc, z_loc, r_loc, top):
        buf = np.zeros(4096, dtype=np.uint64)
        n = lib().oracle_find_bins(self._h, phi_loc, z_loc, r_loc, int(top), _p(buf), buf.size)
        return buf[:n].astype(np.int64)

    def bin_index(self, phi, z, r):
        return int(lib().oracle_bin_index(self._h, phi, z, r))

    def vertex_windows(self, vertex_z, vertex_var_z):
        """The reference's z windows of a vertex list (GridTripletSeedingAlgorithm.cpp:187-206)."""
        vz = np.ascontiguousarray(vertex_z, dtype=np.float64)
        vv = np.ascontiguousarray(vertex_var_z, dtype=np.float64)
        lo = np.zeros(vz.size, np.float32)
        hi = np.zeros(vz.size, np.float32)
        lib().oracle_vertex_windows(C.byref(self._cfg), vz.size, _p(vz), _p(vv), _p(lo), _p(hi))
        return list(zip(lo.tolist(), hi.tolist()))

    def run(self, ev: dict, sort_mode: int = 0, dump_doublets: bool = False,
            z_windows=None, phi_override=None, want_grid: bool = False, vertices=None,
            strip_cot_theta_diff_max=None) -> dict:
        """``strip_cot_theta_diff_max`` given (``float("inf")`` = the reference's default): the strip triplet path
        (TripletSeedFinder.cpp:164-406) with ``ev["strip"]`` = (n, 12) outer-strip calibration details."""
        cols = [np.ascontiguousarray(ev[k], dtype=np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")]
        n = cols[0].size
        if strip_cot_theta_diff_max is not None:
            strip = np.ascontiguousarray(ev["strip"], dtype=np.float32).reshape(n, 12)
            res = C.c_void_p()
            rc = lib().oracle_run_strips(self._h, n, *[_p(c) for c in cols], _p(strip), float(strip_cot_theta_diff_max),
                                         sort_mode, C.byref(res))
            return self._collect(rc, res, False, False)
        if vertices is not None:
            z_windows = self.vertex_windows(*vertices)
        lo = hi = None
        nzw = 0
        if z_windows is not None and len(z_windows) > 0:
            lo = np.ascontiguousarray([w[0] for w in z_windows], dtype=np.float32)
            hi = np.ascontiguousarray([w[1] for w in z_windows], dtype=np.float32)
            nzw = lo.size
        phi = None if phi_override is None else np.ascontiguousarray(phi_override, dtype=np.float32)
        res = C.c_void_p()
        rc = lib().oracle_run(self._h, n, *[_p(c) for c in cols], nzw, _p(lo), _p(hi),
                              sort_mode, int(dump_doublets), _p(phi), C.byref(res))
        return self._collect(rc, res, want_grid, dump_doublets)

    def _collect(self, rc, res, want_grid, dump_doublets) -> dict:
        if rc != 0:
            raise OracleError(rc, lib().oracle_last_error().decode())
        try:
            ns = lib().oracle_result_num_seeds(res)
            out = {
                "bottom": np.zeros(ns, np.uint32), "middle": np.zeros(ns, np.uint32),
                "top": np.zeros(ns, np.uint32), "quality": np.zeros(ns, np.float32),
                "vertexZ": np.zeros(ns, np.float32),
            }
            lib().oracle_result_seeds(res, *[_p(out[k]) for k in ("bottom", "middle", "top", "quality", "vertexZ")])
            cnt = OracleCounters()
            lib().oracle_result_counters(res, C.byref(cnt))
            out["counters"] = cnt.as_dict()
            hist = np.zeros(96, np.uint64)
            lib().oracle_result_histograms(res, _p(hist))
            out["histograms"] = {"bottoms/128": hist[:32].copy(), "tops/128": hist[32:64].copy(), "candPerRound/32": hist[64:].copy()}
            if self._orth is not None:
                order = np.zeros(int(lib().oracle_result_tree_order(res, None)), np.uint32)
                lib().oracle_result_tree_order(res, _p(order))
                out["tree_order"] = order
                ng = lib().oracle_result_grid_size(res)
                core = {"copiedFromIndex": np.zeros(ng, np.uint32), "binBegin": np.zeros(0, np.uint32), "binEnd": np.zeros(0, np.uint32)}
                for k in ("x", "y", "z", "r", "varZ", "varR"):
                    core[k] = np.zeros(ng, np.float32)
                lib().oracle_result_grid(res, *[_p(core[k]) for k in ("copiedFromIndex", "x", "y", "z", "r", "varZ", "varR", "binBegin", "binEnd")])
                out["core"] = core
            elif want_grid or dump_doublets:
                ng = lib().oracle_result_grid_size(res)
                nb = self.info().nGlobalBins
                g = {"copiedFromIndex": np.zeros(ng, np.uint32)}
                for k in ("x", "y", "z", "r", "varZ", "varR"):
                    g[k] = np.zeros(ng, np.float32)
                g["binBegin"] = np.zeros(nb, np.uint32)
                g["binEnd"] = np.zeros(nb, np.uint32)
                lib().oracle_result_grid(res, *[_p(g[k]) for k in ("copiedFromIndex", "x", "y", "z", "r", "varZ", "varR", "binBegin", "binEnd")])
                out["grid"] = g
            if dump_doublets:
                nm = lib().oracle_result_dump_middles(res)
                nd = lib().oracle_result_dump_doublets(res)
                d = {"middlePos": np.zeros(nm, np.uint32), "firstDoublet": np.zeros(nm + 1, np.uint64),
                     "nBottom": np.zeros(nm, np.uint32), "otherPos": np.zeros(nd, np.uint32)}
                for k in ("cotTheta", "iDeltaR", "er", "u", "v", "xNew", "yNew"):
                    d[k] = np.zeros(nd, np.float32)
                lib().oracle_result_dump(res, *[_p(d[k]) for k in ("middlePos", "firstDoublet", "nBottom", "otherPos", "cotTheta", "iDeltaR", "er", "u", "v", "xNew", "yNew")])
                out["doublets"] = d
            return out
        finally:
            lib().oracle_result_free(res)

    def run_many(self, cols: dict, offsets: np.ndarray, n_threads: int = 1, nav_stride: int = 1):
        """Timed-baseline entry: returns per-event seed counts.  ``nav_stride`` > 1
        seeds only every nav_stride-th middle bin (bounded sample of the event)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        arrs = [np.ascontiguousarray(cols[k], dtype=np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")]
        counts = np.zeros(offsets.size - 1, dtype=np.uint64)
        tot = lib().oracle_run_many(self._h, offsets.size - 1, _p(offsets), *[_p(a) for a in arrs], n_threads, nav_stride, _p(counts))
        if tot < 0:
            raise OracleError(-1, "oracle_run_many failed")
        return counts


def estimate_params(seeds: dict, ev: dict, b_field=(0.0, 0.0, 2 * 0.000299792458)) -> np.ndarray:
    """Reference arithmetic of Acts::estimateTrackParamsFromSeed for every seed -> (n, 8) float64."""
    n = int(seeds["bottom"].size)
    idx = [np.ascontiguousarray(seeds[k], dtype=np.uint32) for k in ("bottom", "middle", "top")]
    cols = [np.ascontiguousarray(ev[k], dtype=np.float32) for k in ("x", "y", "z")]
    bf = np.ascontiguousarray(b_field, dtype=np.float64)
    out = np.zeros((n, 8), dtype=np.float64)
    lib().oracle_estimate_params(n, *[_p(a) for a in idx], *[_p(c) for c in cols], _p(bf), _p(out))
    return out


def make_pixel_spacepoints(meas: dict, transforms: np.ndarray) -> dict:
    """Reference arithmetic of SpacePointMaker's createPixelSpacePoint for every measurement.

    ``meas``: surface (uint32), loc0, loc1, cov00, cov01, cov11 (float64); ``transforms``: (nSurfaces, 3, 4)."""
    n = int(meas["surface"].size)
    sf = np.ascontiguousarray(meas["surface"], dtype=np.uint32)
    cols = [np.ascontiguousarray(meas[k], dtype=np.float64) for k in ("loc0", "loc1", "cov00", "cov01", "cov11")]
    tr = np.ascontiguousarray(transforms, dtype=np.float64).reshape(-1, 12)
    out = {k: np.zeros(n, np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")}
    rc = lib().oracle_make_pixel_spacepoints(n, _p(sf), *[_p(c) for c in cols], tr.shape[0], _p(tr),
                                             *[_p(out[k]) for k in ("x", "y", "z", "r", "varZ", "varR")])
    if rc != 0:
        raise ValueError("surface index out of range")
    return out


def seed_set(res: dict) -> dict:
    """{(bottom, middle, top): (quality bits, vertexZ bits)} for exact comparison."""
    q = res["quality"].view(np.uint32)
    z = res["vertexZ"].view(np.uint32)
    return {(int(b), int(m), int(t)): (int(qq), int(zz))
            for b, m, t, qq, zz in zip(res["bottom"], res["middle"], res["top"], q, z)}
