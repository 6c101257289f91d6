"""ctypes wrapper of ``oracle/_ref/libseeding_ref.so``: the UNMODIFIED reference
``GridTripletSeedingAlgorithm`` (see ``oracle/ref_driver.cpp`` and ``oracle/Makefile``).

TEST INFRASTRUCTURE ONLY, like ``oracle/oracle.py``.  The library can only be built
where ``/root/reference`` is mounted (``build()``); the prebuilt ``.so`` travels to
the GPU box with the repository snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from acts_b200.config import Config, OrthogonalOptions

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libseeding_ref.so")
REFERENCE_ROOT = os.environ.get("ACTS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def build(force: bool = False) -> str | None:
    """(Re)build from the reference sources when they are present; otherwise keep the prebuilt file."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "Core", "src", "Seeding")):
        return _LIB_PATH if available() else None
    cmd = ["make", "-C", _HERE, "-j8", "ref", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(_LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.ref_create_orthogonal.argtypes = [C.POINTER(Config), C.POINTER(OrthogonalOptions), C.POINTER(C.c_void_p)]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_run.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.ref_run_strips.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 7 + [C.c_float, C.c_int, C.POINTER(C.c_void_p)]
        L.ref_result_num_seeds.argtypes = [C.c_void_p]
        L.ref_result_num_seeds.restype = C.c_uint64
        L.ref_result_seeds.argtypes = [C.c_void_p] * 6
        L.ref_result_free.argtypes = [C.c_void_p]
        L.ref_run_many.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p]
        L.ref_run_many.restype = C.c_int64
        _lib = L
    return _lib


class ReferenceError_(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"reference error {code}: {msg}")
        self.code = code


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Reference:
    """The reference algorithm object.  ``cfg.useVertexZCuts`` configures ``inputVertices``."""

    def __init__(self, cfg: Config, orthogonal: OrthogonalOptions | None = None):
        """``orthogonal`` given: the reference's OrthogonalTripletSeedingAlgorithm instead."""
        self._h = C.c_void_p()
        self._cfg = cfg  # keeps the arrays the struct points to alive
        self._with_vertices = bool(cfg.useVertexZCuts) and orthogonal is None
        if orthogonal is not None:
            rc = lib().ref_create_orthogonal(C.byref(cfg), C.byref(orthogonal), C.byref(self._h))
        else:
            rc = lib().ref_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise ReferenceError_(rc, lib().ref_last_error().decode())

    def close(self):
        if self._h:
            lib().ref_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, ev: dict, vertices=None) -> dict:
        vertex_z, vertex_var_z = vertices if vertices is not None else (None, None)
        cols = [np.ascontiguousarray(ev[k], dtype=np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")]
        n = cols[0].size
        vz = vv = None
        nv = 0
        if self._with_vertices:
            vz = np.ascontiguousarray(vertex_z if vertex_z is not None else [], dtype=np.float64)
            vv = np.ascontiguousarray(vertex_var_z if vertex_var_z is not None else [], dtype=np.float64)
            nv = vz.size
        res = C.c_void_p()
        rc = lib().ref_run(self._h, n, *[_p(c) for c in cols], nv, _p(vz), _p(vv), C.byref(res))
        return self._collect(rc, res)

    def run_strips(self, ev: dict, cot_theta_diff_max: float = float("inf"), use_strip_info: bool = True) -> dict:
        """The strip triplet path (TripletSeedFinder.cpp:164-406) through ref_run_strips: the reference's Core objects
        driven like execute() drives them, TripletSeedFinder created with ``useStripInfo`` and
        ``cotThetaDiffMax``.  ``ev["strip"]``: (n, 12) float32 outer-strip calibration details.
        ``use_strip_info=False``: the pixel path through the same glue (must equal ``run``)."""
        cols = [np.ascontiguousarray(ev[k], dtype=np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")]
        n = cols[0].size
        strip = np.ascontiguousarray(ev["strip"], dtype=np.float32).reshape(n, 12)
        res = C.c_void_p()
        rc = lib().ref_run_strips(self._h, n, *[_p(c) for c in cols], _p(strip), float(cot_theta_diff_max),
                                  1 if use_strip_info else 0, C.byref(res))
        return self._collect(rc, res)

    @staticmethod
    def _collect(rc, res) -> dict:
        if rc != 0:
            raise ReferenceError_(rc, lib().ref_last_error().decode())
        try:
            ns = lib().ref_result_num_seeds(res)
            out = {
                "bottom": np.zeros(ns, np.uint32), "middle": np.zeros(ns, np.uint32),
                "top": np.zeros(ns, np.uint32), "quality": np.zeros(ns, np.float32),
                "vertexZ": np.zeros(ns, np.float32),
            }
            lib().ref_result_seeds(res, *[_p(out[k]) for k in ("bottom", "middle", "top", "quality", "vertexZ")])
            return out
        finally:
            lib().ref_result_free(res)

    def run_many(self, cols: dict, offsets: np.ndarray, n_threads: int = 1) -> np.ndarray:
        """Timed-baseline entry: whole events, one ``execute`` per event from ``n_threads`` worker threads
        sharing this algorithm object (the Sequencer's pattern).  Returns the per-event seed counts."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        arrs = [np.ascontiguousarray(cols[k], dtype=np.float32) for k in ("x", "y", "z", "r", "varZ", "varR")]
        counts = np.zeros(offsets.size - 1, dtype=np.uint64)
        tot = lib().ref_run_many(self._h, offsets.size - 1, _p(offsets), *[_p(a) for a in arrs], int(n_threads), _p(counts))
        if tot < 0:
            raise ReferenceError_(-1, lib().ref_last_error().decode())
        return counts
