"""ctypes binding of the C ABI in ``include/acts_b200_seeding.h``.

This is what a host application does across the FFI boundary: fill a
``b200seed_config`` (same fields as ``GridTripletSeedingAlgorithm::Config``),
create a handle, pass raw column pointers, receive seed columns.  There is no
CPU fallback: if the CUDA library is missing or no device is present the calls
raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import config as cfgmod
from .config import Config, Counters, Doublets, Info, OrthogonalOptions, Seeds

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SEED_LIB", os.path.join(_HERE, "libacts_b200_seeding.so"))  # override: kernel experiments

# every symbol include/acts_b200_seeding.h declares
EXPORTED_SYMBOLS = [
    "b200seed_config_init", "b200seed_plan_info", "b200seed_plan_tables", "b200seed_create",
    "b200seed_orthogonal_config_init", "b200seed_create_orthogonal",
    "b200seed_destroy", "b200seed_last_error", "b200seed_alloc_pinned", "b200seed_free_pinned",
    "b200seed_get_info", "b200seed_get_counters",
    "b200seed_get_stage_times", "b200seed_get_stage_times_ex", "b200seed_set_phi_sector", "b200seed_estimate_params",
    "b200seed_run_vertices", "b200seed_vertex_windows", "b200seed_run_batch_windows", "b200seed_run_strips",
    "b200seed_make_pixel_spacepoints", "b200seed_run_measurements",
    "b200seed_run", "b200seed_run_with_phi", "b200seed_run_batch", "b200seed_run_batch_device",
    "b200seed_sync", "b200seed_debug_grid", "b200seed_debug_doublets", "b200seed_debug_atan2f",
]


class SeedingError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b200seed error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def lib():
    """Load the plugin (built in-tree by ``acts_b200.build``); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SeedingError(cfgmod.ERR_CUDA, f"{LIB_PATH} is missing: run `python -m acts_b200.build` (nvcc, sm_100a)")
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, f32p = C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p
        L.b200seed_last_error.restype = C.c_char_p
        L.b200seed_alloc_pinned.restype = vp
        L.b200seed_alloc_pinned.argtypes = [C.c_size_t]
        L.b200seed_free_pinned.argtypes = [vp]
        L.b200seed_config_init.argtypes = [C.POINTER(Config)]
        L.b200seed_plan_info.argtypes = [C.POINTER(Config), C.POINTER(Info)]
        L.b200seed_plan_tables.argtypes = [C.POINTER(Config), vp, u64, vp, vp, vp, vp, vp, vp]
        L.b200seed_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(vp)]
        L.b200seed_orthogonal_config_init.argtypes = [C.POINTER(Config), C.POINTER(OrthogonalOptions)]
        L.b200seed_create_orthogonal.argtypes = [C.POINTER(Config), C.POINTER(OrthogonalOptions), C.c_int, C.POINTER(vp)]
        L.b200seed_destroy.argtypes = [vp]
        L.b200seed_get_info.argtypes = [vp, C.POINTER(Info)]
        L.b200seed_get_counters.argtypes = [vp, C.POINTER(Counters)]
        L.b200seed_get_stage_times.argtypes = [vp, vp]
        L.b200seed_get_stage_times_ex.argtypes = [vp, vp, u32]
        L.b200seed_run_vertices.argtypes = [vp, u32] + [f32p] * 6 + [u32, vp, vp, C.POINTER(Seeds)]
        L.b200seed_run_strips.argtypes = [vp, u32] + [f32p] * 7 + [C.c_float, C.POINTER(Seeds)]
        L.b200seed_vertex_windows.argtypes = [vp, u32, vp, vp, vp, vp]
        L.b200seed_run_batch_windows.argtypes = [vp, u32, vp] + [f32p] * 6 + [vp, f32p, f32p, vp, C.POINTER(Seeds)]
        L.b200seed_set_phi_sector.argtypes = [vp, u32, u32]
        L.b200seed_estimate_params.argtypes = [vp, u64, vp, vp, vp, u32, vp, vp, vp, vp, vp]
        L.b200seed_make_pixel_spacepoints.argtypes = [vp, u32] + [vp] * 6 + [u32] + [vp] * 7
        L.b200seed_run_measurements.argtypes = [vp, u32] + [vp] * 6 + [u32, vp, u32, vp, vp] + [vp] * 6 + [C.POINTER(Seeds)]
        L.b200seed_run.argtypes = [vp, u32] + [f32p] * 6 + [u32, f32p, f32p, C.POINTER(Seeds)]
        L.b200seed_run_with_phi.argtypes = [vp, u32] + [f32p] * 7 + [C.POINTER(Seeds)]
        L.b200seed_run_batch.argtypes = [vp, u32, vp] + [f32p] * 6 + [vp, C.POINTER(Seeds)]
        L.b200seed_run_batch_device.argtypes = [vp, u32, u32, vp] + [f32p] * 6 + [vp, C.POINTER(Seeds), vp]
        L.b200seed_sync.argtypes = [vp, C.POINTER(Seeds)]
        L.b200seed_debug_grid.argtypes = [vp, u64] + [vp] * 7 + [u64, vp, vp]
        L.b200seed_debug_doublets.argtypes = [vp, C.POINTER(Doublets)]
        L.b200seed_debug_atan2f.argtypes = [vp, u64, vp, vp, vp]
        _lib = L
    return _lib


def config_init(cfg_ref):
    lib().b200seed_config_init(cfg_ref)


def orthogonal_config_init(cfg_ref, opt_ref):
    lib().b200seed_orthogonal_config_init(cfg_ref, opt_ref)


def _check(rc: int):
    if rc != 0:
        raise SeedingError(rc, lib().b200seed_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def plan_info(cfg: Config) -> Info:
    """Host-only validation + derived constants (no GPU needed)."""
    info = Info()
    _check(lib().b200seed_plan_info(C.byref(cfg), C.byref(info)))
    return info


def plan_tables(cfg: Config) -> dict:
    """Host-only: device constant block + navigation / neighbour tables."""
    sizes = np.zeros(8, dtype=np.uint64)
    _check(lib().b200seed_plan_tables(C.byref(cfg), None, 0, None, None, None, None, None, _p(sizes)))
    n_nav, n_bot, n_top, dc_bytes, k = (int(v) for v in sizes[:5])
    out = {
        "deviceConfig": np.zeros(dc_bytes, dtype=np.uint8),
        "navBins": np.zeros(n_nav, dtype=np.uint32),
        "botOffsets": np.zeros(n_nav + 1, dtype=np.uint32),
        "botBins": np.zeros(max(n_bot, 1), dtype=np.uint32),
        "topOffsets": np.zeros(n_nav + 1, dtype=np.uint32),
        "topBins": np.zeros(max(n_top, 1), dtype=np.uint32),
        "seedsPerMiddle": k,
    }
    _check(lib().b200seed_plan_tables(C.byref(cfg), _p(out["deviceConfig"]), dc_bytes, _p(out["navBins"]),
                                      _p(out["botOffsets"]), _p(out["botBins"]), _p(out["topOffsets"]),
                                      _p(out["topBins"]), _p(sizes)))
    out["botBins"] = out["botBins"][:n_bot]
    out["topBins"] = out["topBins"][:n_top]
    return out


class SeedingEngine:
    """One handle = one device + one stream (see the header's threading contract)."""

    def __init__(self, cfg: Config, device: int = 0, orthogonal: OrthogonalOptions | None = None):
        """``orthogonal`` given: an OrthogonalTripletSeedingAlgorithm handle (k-d-tree candidate provider)."""
        self._h = C.c_void_p()
        self.cfg = cfg
        self.orthogonal = orthogonal
        if orthogonal is not None:
            _check(lib().b200seed_create_orthogonal(C.byref(cfg), C.byref(orthogonal), device, C.byref(self._h)))
        else:
            _check(lib().b200seed_create(C.byref(cfg), device, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().b200seed_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> Info:
        i = Info()
        _check(lib().b200seed_get_info(self._h, C.byref(i)))
        return i

    def counters(self) -> dict:
        c = Counters()
        _check(lib().b200seed_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def set_phi_sector(self, first_phi_bin: int = 1, n_phi_bins: int = 0, empty: bool = False):
        """Seed only middles in phi bins [first, first + n) (1-based); n = 0 -> all; ``empty``: no bin at all
        (a rank of a split with more ranks than phi bins)."""
        if empty:
            first_phi_bin, n_phi_bins = 0xFFFFFFFF, 0
        _check(lib().b200seed_set_phi_sector(self._h, first_phi_bin, n_phi_bins))

    def estimate_params(self, seeds: dict, ev: dict, b_field=(0.0, 0.0, 2 * 0.000299792458)) -> np.ndarray:
        """Free parameters (n, 8) of the seeds: ``b200seed_estimate_params``."""
        n = int(seeds["bottom"].size)
        idx = [np.ascontiguousarray(seeds[k], dtype=np.uint32) for k in ("bottom", "middle", "top")]
        cols = [np.ascontiguousarray(ev[k], dtype=np.float32) for k in ("x", "y", "z")]
        bf = np.ascontiguousarray(b_field, dtype=np.float64)
        out = np.zeros((n, 8), dtype=np.float64)
        _check(lib().b200seed_estimate_params(self._h, n, *[_p(a) for a in idx], cols[0].size, *[_p(c) for c in cols], _p(bf), _p(out)))
        return out

    def make_pixel_spacepoints(self, meas: dict, transforms: np.ndarray) -> dict:
        """Measurements on planar surfaces -> the six space-point columns: ``b200seed_make_pixel_spacepoints``.

        ``meas``: surface (uint32), loc0, loc1, cov00, cov01, cov11 (float64); ``transforms``: (nSurfaces, 3, 4)."""
        n = int(meas["surface"].size)
        sf = np.ascontiguousarray(meas["surface"], dtype=np.uint32)
        cols = [np.ascontiguousarray(meas[k], dtype=np.float64) for k in ("loc0", "loc1", "cov00", "cov01", "cov11")]
        tr = np.ascontiguousarray(transforms, dtype=np.float64).reshape(-1, 12)
        out = {k: np.zeros(n, np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")}
        _check(lib().b200seed_make_pixel_spacepoints(self._h, n, _p(sf), *[_p(c) for c in cols], tr.shape[0], _p(tr),
                                                     *[_p(out[k]) for k in ("x", "y", "z", "r", "varZ", "varR")]))
        return out

    def run_measurements(self, meas: dict, transforms: np.ndarray, want_spacepoints: bool = False):
        """Measurements -> space points -> seeds in one call (``b200seed_run_measurements``)."""
        n = int(meas["surface"].size)
        sf = np.ascontiguousarray(meas["surface"], dtype=np.uint32)
        cols = [np.ascontiguousarray(meas[k], dtype=np.float64) for k in ("loc0", "loc1", "cov00", "cov01", "cov11")]
        tr = np.ascontiguousarray(transforms, dtype=np.float64).reshape(-1, 12)
        sp = {k: np.zeros(n, np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")} if want_spacepoints else None
        out, s = self._alloc(max(16, n * 6))
        sp_ptrs = [_p(sp[k]) for k in ("x", "y", "z", "r", "varZ", "varR")] if sp is not None else [None] * 6
        _check(lib().b200seed_run_measurements(self._h, n, _p(sf), *[_p(c) for c in cols], tr.shape[0], _p(tr), 0, None, None,
                                               *sp_ptrs, C.byref(s)))
        seeds = {name: arr[:int(s.size)] for name, arr in out.items()}
        return (seeds, sp) if want_spacepoints else seeds

    def stage_times_ms(self) -> dict:
        ms = np.zeros(7, dtype=np.float32)
        _check(lib().b200seed_get_stage_times_ex(self._h, _p(ms), 7))
        return {"grid": float(ms[0]), "work": float(ms[1]), "seed": float(ms[2]), "compact": float(ms[3]),
                "doublet_count": float(ms[4]), "doublet_fill": float(ms[5]), "seed_middles": float(ms[6])}

    def vertex_windows(self, vertex_z, vertex_var_z):
        """The reference's z windows of a vertex list (``b200seed_vertex_windows``)."""
        vz = np.ascontiguousarray(vertex_z, dtype=np.float64)
        vv = np.ascontiguousarray(vertex_var_z, dtype=np.float64)
        lo, hi = np.zeros(vz.size, np.float32), np.zeros(vz.size, np.float32)
        _check(lib().b200seed_vertex_windows(self._h, vz.size, _p(vz), _p(vv), _p(lo), _p(hi)))
        return lo, hi

    @staticmethod
    def _cols(ev):
        return [np.ascontiguousarray(ev[k], dtype=np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")]

    @staticmethod
    def _alloc(capacity):
        out = {
            "bottom": np.zeros(capacity, np.uint32), "middle": np.zeros(capacity, np.uint32),
            "top": np.zeros(capacity, np.uint32), "quality": np.zeros(capacity, np.float32),
            "vertexZ": np.zeros(capacity, np.float32),
        }
        s = Seeds()
        s.bottom, s.middle, s.top = _p(out["bottom"]), _p(out["middle"]), _p(out["top"])
        s.quality, s.vertexZ = _p(out["quality"]), _p(out["vertexZ"])
        s.capacity = capacity
        return out, s

    def run(self, ev: dict, z_windows=None, phi=None, capacity=None, out=None, vertices=None,
            strip_cot_theta_diff_max=None) -> dict:
        """One event through ``b200seed_run`` (host buffers in, host seeds out).

        ``vertices`` = (z, var z) goes through ``b200seed_run_vertices`` (Config::inputVertices).
        ``strip_cot_theta_diff_max`` given: ``b200seed_run_strips`` with ``ev["strip"]`` ((n, 12) float32).
        ``out`` may hold caller-owned (e.g. pinned) seed columns that are reused from call to call."""
        cols = self._cols(ev)
        n = cols[0].size
        cap = capacity if capacity is not None else max(16, n * (12 if self.orthogonal is not None else 6))
        if out is not None:
            s = Seeds()
            s.bottom, s.middle, s.top = _p(out["bottom"]), _p(out["middle"]), _p(out["top"])
            s.quality, s.vertexZ = _p(out["quality"]), _p(out["vertexZ"])
            s.capacity = min(int(a.size) for a in out.values())
        else:
            out, s = self._alloc(cap)
        if strip_cot_theta_diff_max is not None:
            strip = np.ascontiguousarray(ev["strip"], dtype=np.float32).reshape(n, 12)
            rc = lib().b200seed_run_strips(self._h, n, *[_p(c) for c in cols], _p(strip), float(strip_cot_theta_diff_max), C.byref(s))
        elif vertices is not None:
            vz = np.ascontiguousarray(vertices[0], dtype=np.float64)
            vv = np.ascontiguousarray(vertices[1], dtype=np.float64)
            rc = lib().b200seed_run_vertices(self._h, n, *[_p(c) for c in cols], vz.size, _p(vz), _p(vv), C.byref(s))
        elif phi is not None:
            phi = np.ascontiguousarray(phi, dtype=np.float32)
            rc = lib().b200seed_run_with_phi(self._h, n, *[_p(c) for c in cols], _p(phi), C.byref(s))
        else:
            lo = hi = None
            nzw = 0
            if z_windows is not None and len(z_windows) > 0:
                lo = np.ascontiguousarray([w[0] for w in z_windows], dtype=np.float32)
                hi = np.ascontiguousarray([w[1] for w in z_windows], dtype=np.float32)
                nzw = lo.size
            rc = lib().b200seed_run(self._h, n, *[_p(c) for c in cols], nzw, _p(lo), _p(hi), C.byref(s))
        _check(rc)
        k = int(s.size)
        return {name: arr[:k] for name, arr in out.items()}

    def run_batch(self, cols: dict, offsets: np.ndarray, capacity=None, out=None, z_windows=None) -> list[dict]:
        """A batch of events through ``b200seed_run_batch``; returns one dict per event.

        ``z_windows``: one list of (lo, hi) per event -> ``b200seed_run_batch_windows``.

        ``out`` may hold caller-owned (e.g. pinned) seed columns ``bottom/middle/top`` (uint32) and
        ``quality/vertexZ`` (float32) that are reused from call to call."""
        arrs = self._cols(cols)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        n_events = offsets.size - 1
        cap = capacity if capacity is not None else max(16, int(offsets[-1]) * (12 if self.orthogonal is not None else 6))
        if out is not None:
            cap = min(int(a.size) for a in out.values())
            s = Seeds()
            s.bottom, s.middle, s.top = _p(out["bottom"]), _p(out["middle"]), _p(out["top"])
            s.quality, s.vertexZ = _p(out["quality"]), _p(out["vertexZ"])
            s.capacity = cap
        else:
            out, s = self._alloc(cap)
        seed_offsets = np.zeros(n_events + 1, dtype=np.uint64)
        if z_windows is not None:
            w_off = np.zeros(n_events + 1, dtype=np.uint32)
            w_off[1:] = np.cumsum([len(w) for w in z_windows])
            lo = np.ascontiguousarray([w[0] for ws in z_windows for w in ws], dtype=np.float32)
            hi = np.ascontiguousarray([w[1] for ws in z_windows for w in ws], dtype=np.float32)
            _check(lib().b200seed_run_batch_windows(self._h, n_events, _p(offsets), *[_p(a) for a in arrs], _p(w_off), _p(lo), _p(hi),
                                                    _p(seed_offsets), C.byref(s)))
        else:
            _check(lib().b200seed_run_batch(self._h, n_events, _p(offsets), *[_p(a) for a in arrs], _p(seed_offsets), C.byref(s)))
        res = []
        for e in range(n_events):
            a, b = int(seed_offsets[e]), int(seed_offsets[e + 1])
            res.append({name: arr[a:b] for name, arr in out.items()})
        return res

    def debug_grid(self, n_events: int = 1) -> dict:
        n = self.counters()["nInGrid"]
        nb = self.info().nGlobalBins * n_events
        g = {"copiedFromIndex": np.zeros(n, np.uint32)}
        for k in ("x", "y", "z", "r", "varZ", "varR"):
            g[k] = np.zeros(n, np.float32)
        g["binBegin"] = np.zeros(nb, np.uint32)
        g["binEnd"] = np.zeros(nb, np.uint32)
        _check(lib().b200seed_debug_grid(self._h, n, *[_p(g[k]) for k in ("copiedFromIndex", "x", "y", "z", "r", "varZ", "varR")],
                                         nb, _p(g["binBegin"]), _p(g["binEnd"])))
        return g

    def debug_doublets(self, fetch: bool = True) -> dict:
        """Materialised two-pass doublet search of the last batch (``b200seed_debug_doublets``)."""
        d = Doublets()
        _check(lib().b200seed_debug_doublets(self._h, C.byref(d)))
        nm, nd = int(d.nMiddles), int(d.nDoublets)
        out = {"nMiddles": nm, "nDoublets": nd}
        if not fetch:
            return out
        arr = {"middlePos": np.zeros(nm, np.uint32), "firstDoublet": np.zeros(nm + 1, np.uint64),
               "nBottom": np.zeros(nm, np.uint32), "otherPos": np.zeros(nd, np.uint32)}
        for k in ("cotTheta", "iDeltaR", "er", "u", "v", "xNew", "yNew"):
            arr[k] = np.zeros(nd, np.float32)
        for k, a in arr.items():
            setattr(d, k, _p(a))
        d.middleCapacity, d.doubletCapacity = nm, nd
        _check(lib().b200seed_debug_doublets(self._h, C.byref(d)))
        out.update(arr)
        out["gpuMilliseconds"] = float(d.gpuMilliseconds)
        return out

    def device_atan2f(self, y: np.ndarray, x: np.ndarray) -> np.ndarray:
        y = np.ascontiguousarray(y, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(y)
        _check(lib().b200seed_debug_atan2f(self._h, y.size, _p(y), _p(x), _p(out)))
        return out

    # ---- device-resident path (inputs already in HBM) -------------------------
    def run_batch_device(self, n_events, n_total, d_offsets, d_cols, d_seed_offsets, d_out, capacity, stream=None):
        """Raw device pointers (ints); asynchronous, see ``b200seed_run_batch_device``."""
        s = Seeds()
        s.bottom, s.middle, s.top, s.quality, s.vertexZ = d_out
        s.capacity = capacity
        _check(lib().b200seed_run_batch_device(self._h, n_events, n_total, d_offsets, *d_cols, d_seed_offsets, C.byref(s), stream))
        return s

    def sync(self, seeds: Seeds | None = None) -> int:
        s = seeds if seeds is not None else Seeds()
        _check(lib().b200seed_sync(self._h, C.byref(s)))
        return int(s.size)
