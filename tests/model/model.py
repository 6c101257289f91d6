"""ctypes wrapper of the sequential CPU model of the GPU algorithm (test infra)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from acts_b200 import build, plugin

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build.build_model())
        vp = C.c_void_p
        L.model_run.argtypes = [vp] * 10 + [C.c_uint32] + [vp] * 7 + [C.c_uint32, vp, vp, C.c_int, C.c_uint64] + [vp] * 6
        L.model_run.restype = C.c_int64
        L.model_bin_index.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float]
        L.model_bin_index.restype = C.c_int32
        L.model_atan2f.argtypes = [C.c_float, C.c_float]
        L.model_atan2f.restype = C.c_float
        L.model_check_std_sort.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.model_check_std_sort.restype = C.c_int64
        L.model_check_std_sort_killer.argtypes = [C.c_int]
        L.model_check_std_sort_killer.restype = C.c_int64
        L.model_check_tie_replay.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.model_check_tie_replay.restype = C.c_int64
        L.model_check_warp_replay.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]
        L.model_check_warp_replay.restype = C.c_int64
        L.model_check_heap.argtypes = [C.c_uint64, C.c_int]
        L.model_check_heap.restype = C.c_int64
        L.model_check_atan2f.argtypes = [C.c_uint64, C.c_int64]
        L.model_check_atan2f.restype = C.c_int64
        L.model_sizeof_device_config.restype = C.c_uint64
        L.model_set_strips.argtypes = [vp, C.c_float, C.c_float]
        L.model_set_strips.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def radius_ranges(cfg, tables, grid):
    """Per navigation entry (rMin, rMax) exactly like
    GridTripletSeedingAlgorithm::retrieveRadiusRangeForMiddle (.cpp:404-421)."""
    nav = tables["navBins"]
    lo = np.zeros(nav.size, np.float32)
    hi = np.zeros(nav.size, np.float32)
    if cfg.useVariableMiddleSPRange:
        mins = [grid["r"][b] for b, e in zip(grid["binBegin"], grid["binEnd"]) if b != e]
        maxs = [grid["r"][e - 1] for b, e in zip(grid["binBegin"], grid["binEnd"]) if b != e]
        mn = np.float32(min(mins)) if mins else np.float32(np.finfo(np.float32).max)
        mx = np.float32(max(maxs)) if maxs else np.float32(np.finfo(np.float32).min)
        lo[:] = np.float32(np.floor(mn / np.float32(2)) * np.float32(2)) + np.float32(cfg.deltaRMiddleMinSPRange)
        hi[:] = np.float32(np.floor(mx / np.float32(2)) * np.float32(2)) - np.float32(cfg.deltaRMiddleMaxSPRange)
        return lo, hi
    if cfg.nRRangeMiddleSP == 0:
        lo[:] = cfg.rMinMiddle
        hi[:] = cfg.rMaxMiddle
        return lo, hi
    edges = np.array([cfg.zBinEdges[i] for i in range(cfg.nZBinEdges)], dtype=np.float32)
    for g, b in enumerate(nav):
        if grid["binBegin"][b] == grid["binEnd"][b]:
            continue
        z_first = grid["z"][grid["binBegin"][b]]
        zb = int(np.searchsorted(edges, z_first, side="left"))
        if zb != 0:
            zb -= 1
        lo[g] = cfg.rRangeMiddleSP[2 * zb]
        hi[g] = cfg.rRangeMiddleSP[2 * zb + 1]
    return lo, hi


def run(cfg, grid: dict, tie_mode: int = 0, z_windows=None, strip=None, cot_theta_diff_max=float("inf")) -> dict:
    """Run the restructured algorithm on a packed grid (dict like Oracle.run()['grid']).
    ``strip`` ((n, 12) float32, per ORIGINAL space point): the strip triplet path with ``cot_theta_diff_max``."""
    t = plugin.plan_tables(cfg)
    dc = t["deviceConfig"].copy()
    assert dc.size == lib().model_sizeof_device_config()
    if z_windows:
        # doubletCuts = kCutsVertexZ: patch like the plugin does per event
        raise NotImplementedError
    lo, hi = radius_ranges(cfg, t, grid)
    cap = max(16, grid["x"].size * 6)
    ob, om, ot = (np.zeros(cap, np.uint32) for _ in range(3))
    oq, oz = (np.zeros(cap, np.float32) for _ in range(2))
    stats = np.zeros(8, np.uint64)
    if strip is not None:
        strip = np.ascontiguousarray(strip, dtype=np.float32)
        lib().model_set_strips(_p(strip), float(cot_theta_diff_max), float(cfg.toleranceParam))
    n = lib().model_run(_p(dc), _p(grid["copiedFromIndex"]), _p(grid["x"]), _p(grid["y"]), _p(grid["z"]), _p(grid["r"]),
                        _p(grid["varZ"]), _p(grid["varR"]), _p(grid["binBegin"]), _p(grid["binEnd"]),
                        t["navBins"].size, _p(t["navBins"]), _p(lo), _p(hi), _p(t["botOffsets"]), _p(t["botBins"]),
                        _p(t["topOffsets"]), _p(t["topBins"]), 0, None, None, tie_mode, cap,
                        _p(ob), _p(om), _p(ot), _p(oq), _p(oz), _p(stats))
    if strip is not None:
        lib().model_set_strips(None, 0.0, 1.1)
    assert 0 <= n <= cap
    return {"bottom": ob[:n], "middle": om[:n], "top": ot[:n], "quality": oq[:n], "vertexZ": oz[:n],
            "stats": {"nBottomDoublets": int(stats[0]), "nTopDoublets": int(stats[1]),
                      "nTripletTests": int(stats[2]), "nCandidates": int(stats[3]),
                      "nConfirmationRounds": int(stats[4])}}


def bin_index(cfg, x, y, z, r):
    t = plugin.plan_tables(cfg)
    dc = t["deviceConfig"]
    return np.array([lib().model_bin_index(_p(dc), float(a), float(b), float(c), float(d)) for a, b, c, d in zip(x, y, z, r)], dtype=np.int64)
