"""ACTS Examples CSV formats at the boundary of the seeding path.

Lets dumps of a real ACTS build (``CsvSpacePointWriter`` / ``CsvSeedWriter`` behind
``seeding.py``) be replayed through the engine and compared with its output.
Host-side file parsing only; nothing here computes what the engine computes.

Formats (reference ``file:line``):

* space points, writer layout ``SpacePointData2``
  (Examples/Io/Csv/src/CsvOutputData.hpp:385-393, written by
  CsvSpacePointWriter.cpp:52-82 as ``event%09d-spacepoint.csv``):
  ``measurement_id_1,measurement_id_2,geometry_id_1,geometry_id_2,x,y,z,t,var_r,var_z``
* space points, reader layout ``SpacePointData``
  (CsvOutputData.hpp:347-383, read by CsvSpacePointReader.cpp:57-75):
  ``measurement_id,sp_x,sp_y,sp_z,sp_radius,sp_covr,sp_covz,...strip columns``
* seeds (CsvSeedWriter.cpp:168-200):
  ``seed_id,particleId,pT,eta,phi,bX,bY,bZ,mX,mY,mZ,tX,tY,tZ,good/duplicate/fake,vertexZ,quality,Hits_ID``
  with ``Hits_ID`` = ``"[id0,id1,id2,]"`` (the measurement ids of bottom, middle, top).

File names follow ``perEventFilepath`` (Examples/Framework/src/Utilities/Paths.cpp:42-54).
"""
from __future__ import annotations

import csv
import os

import numpy as np

SPACEPOINT_WRITER_COLUMNS = ("measurement_id_1", "measurement_id_2", "geometry_id_1", "geometry_id_2",
                             "x", "y", "z", "t", "var_r", "var_z")
SEED_COLUMNS = ("seed_id", "particleId", "pT", "eta", "phi", "bX", "bY", "bZ", "mX", "mY", "mZ", "tX", "tY", "tZ",
                "good/duplicate/fake", "vertexZ", "quality", "Hits_ID")
NO_SECOND_MEASUREMENT = np.uint64(0xFFFFFFFFFFFFFFFF)  # CsvSpacePointWriter.cpp:70


def per_event_filepath(directory: str, name: str, event: int) -> str:
    """``ActsExamples::perEventFilepath`` (Paths.cpp:42-54)."""
    fn = "event%09d-%s" % (event, name)
    return os.path.join(directory, fn) if directory else fn


def _f32(v: float) -> str:
    # max_digits10 of float = 9 significant digits (CsvSpacePointWriter.hpp:36): exact round trip
    return "%.9g" % float(np.float32(v))


def read_spacepoints(path: str) -> dict:
    """Read one event of space points in either layout.

    Returns the six float32 columns of the C ABI plus ``measurement_id`` (uint64).
    ``r`` is ``fastHypot(x, y)`` like SpacePointMaker.cpp:72,161 — evaluated in double on
    the float32 coordinates of the file and rounded to float32 (the reader layout's
    own ``sp_radius`` column is used when present).
    """
    with open(path, newline="") as fh:
        reader = csv.reader(fh)
        header = next(reader)
        rows = [row for row in reader if row]
    col = {name: i for i, name in enumerate(header)}
    n = len(rows)

    def column(name, dtype):
        i = col[name]
        return np.array([row[i] for row in rows], dtype=dtype) if n else np.zeros(0, dtype)

    if "sp_x" in col:  # reader layout
        x, y, z = (column(k, np.float32) for k in ("sp_x", "sp_y", "sp_z"))
        var_r, var_z = column("sp_covr", np.float32), column("sp_covz", np.float32)
        mid = column("measurement_id", np.uint64)
        r = column("sp_radius", np.float32) if "sp_radius" in col else None
    elif "x" in col:  # writer layout
        x, y, z = (column(k, np.float32) for k in ("x", "y", "z"))
        var_r, var_z = column("var_r", np.float32), column("var_z", np.float32)
        mid = column("measurement_id_1", np.uint64)
        r = None
    else:
        raise ValueError(f"{path}: neither the SpacePointData nor the SpacePointData2 layout")
    if r is None or (n and not np.any(r)):
        r = np.sqrt(x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2).astype(np.float32)
    out = {"x": x, "y": y, "z": z, "r": r, "varZ": var_z, "varR": var_r, "measurement_id": mid}
    if "sp_topStripCenterPosition_0" in col:
        # the strip columns of the reader layout -> the StripCalibrationDetails column exactly like
        # CsvSpacePointReader.cpp:80-110 (extendCollection): float columns, the half vectors formed in double
        # (Acts::Vector3 * float) and narrowed to float
        def vec(prefix):
            return np.stack([column("%s_%d" % (prefix, k), np.float32) for k in range(3)], axis=1).astype(np.float64) \
                if n else np.zeros((0, 3), np.float64)

        top_len = column("sp_topHalfStripLength", np.float32).astype(np.float64)
        bot_len = column("sp_bottomHalfStripLength", np.float32).astype(np.float64)
        outer_center = vec("sp_topStripCenterPosition")
        separation = vec("sp_stripCenterDistance")
        outer_half = vec("sp_topStripDirection") * top_len[:, None]
        inner_half = vec("sp_bottomStripDirection") * bot_len[:, None]
        out["strip"] = np.ascontiguousarray(np.concatenate([outer_center, separation, outer_half, inner_half], axis=1), dtype=np.float32)
    return out


STRIP_READER_COLUMNS = ("measurement_id", "sp_x", "sp_y", "sp_z", "sp_radius", "sp_covr", "sp_covz",
                        "sp_topHalfStripLength", "sp_bottomHalfStripLength",
                        "sp_topStripDirection_0", "sp_topStripDirection_1", "sp_topStripDirection_2",
                        "sp_bottomStripDirection_0", "sp_bottomStripDirection_1", "sp_bottomStripDirection_2",
                        "sp_stripCenterDistance_0", "sp_stripCenterDistance_1", "sp_stripCenterDistance_2",
                        "sp_topStripCenterPosition_0", "sp_topStripCenterPosition_1", "sp_topStripCenterPosition_2")


def write_strip_spacepoints(path: str, sp: dict, measurement_id=None) -> None:
    """Write one event of strip space points in the reader layout ``SpacePointData`` (CsvOutputData.hpp:347-383):
    ``sp["strip"]`` ((n, 12): outerCenter, innerToOuterSeparation, outerHalfVector, innerHalfVector) goes out as
    half lengths + unit directions + separation + outer centre, the columns CsvSpacePointReader.cpp:80-97 reads."""
    n = sp["x"].size
    ids = np.arange(n, dtype=np.uint64) if measurement_id is None else np.asarray(measurement_id, np.uint64)
    d = np.asarray(sp["strip"], np.float32).reshape(n, 12).astype(np.float64)
    with open(path, "w", newline="") as fh:
        fh.write(",".join(STRIP_READER_COLUMNS) + "\n")
        for i in range(n):
            oc, sep, oh, ih = d[i, 0:3], d[i, 3:6], d[i, 6:9], d[i, 9:12]
            lt, lb = float(np.linalg.norm(oh)), float(np.linalg.norm(ih))
            dt = oh / lt if lt > 0 else oh
            db = ih / lb if lb > 0 else ih
            vals = [sp["x"][i], sp["y"][i], sp["z"][i], sp["r"][i], sp["varR"][i], sp["varZ"][i], lt, lb, *dt, *db, *sep, *oc]
            fh.write("%d,%s\n" % (int(ids[i]), ",".join(_f32(v) for v in vals)))


def write_spacepoints(path: str, sp: dict, measurement_id=None) -> None:
    """Write one event in the ``CsvSpacePointWriter`` layout (pixel space points: one source link)."""
    n = sp["x"].size
    ids = np.arange(n, dtype=np.uint64) if measurement_id is None else np.asarray(measurement_id, np.uint64)
    with open(path, "w", newline="") as fh:
        fh.write(",".join(SPACEPOINT_WRITER_COLUMNS) + "\n")
        for i in range(n):
            fh.write("%d,%d,0,0,%s,%s,%s,0,%s,%s\n" % (
                int(ids[i]), int(NO_SECOND_MEASUREMENT), _f32(sp["x"][i]), _f32(sp["y"][i]), _f32(sp["z"][i]),
                _f32(sp["varR"][i]), _f32(sp["varZ"][i])))


def write_seeds(path: str, seeds: dict, sp: dict, free_params=None, measurement_id=None) -> None:
    """Write seeds in the ``CsvSeedWriter`` layout.

    ``free_params`` (n x 8, from ``SeedingEngine.estimate_params``) fills pT / eta / phi
    like CsvSeedWriter.cpp:85-87,147-148; without it the three columns are -1, 0, 0
    (the struct defaults, .cpp:36-38).  Truth columns: ``particleId`` 0, type ``unknown``.
    """
    n = seeds["quality"].size
    ids = np.arange(sp["x"].size, dtype=np.uint64) if measurement_id is None else np.asarray(measurement_id, np.uint64)
    with open(path, "w", newline="") as fh:
        fh.write(",".join(SEED_COLUMNS) + "\n")
        for i in range(n):
            b, m, t = int(seeds["bottom"][i]), int(seeds["middle"][i]), int(seeds["top"][i])
            pt, eta, phi = -1.0, 0.0, 0.0
            if free_params is not None:
                d = free_params[i, 4:7]
                qop = free_params[i, 7]
                theta = np.arctan2(np.hypot(d[0], d[1]), d[2])
                phi = float(np.float32(np.arctan2(d[1], d[0])))
                eta = float(np.float32(np.arctanh(np.cos(theta))))
                pt = float(np.float32(abs(1.0 / qop) * np.sin(theta)))
            pos = ",".join("%s,%s,%s" % (_f32(sp["x"][k]), _f32(sp["y"][k]), _f32(sp["z"][k])) for k in (b, m, t))
            fh.write("%d,0,%s,%s,%s,%s,unknown,%s,%s,\"[%d,%d,%d,]\"\n" % (
                i, _f32(pt), _f32(eta), _f32(phi), pos, _f32(seeds["vertexZ"][i]), _f32(seeds["quality"][i]),
                int(ids[b]), int(ids[m]), int(ids[t])))


def read_seeds(path: str, measurement_id=None) -> dict:
    """Read a ``CsvSeedWriter`` file.  ``bottom/middle/top`` are the measurement ids of
    ``Hits_ID`` mapped back to space-point indices when ``measurement_id`` (the column
    returned by ``read_spacepoints``) is given."""
    out = {k: [] for k in ("seed_id", "bottom", "middle", "top", "vertexZ", "quality", "pT", "eta", "phi")}
    with open(path, newline="") as fh:
        reader = csv.DictReader(fh)
        for row in reader:
            hits = [int(v) for v in row["Hits_ID"].strip("[]\" ").split(",") if v.strip()]
            if len(hits) != 3:
                raise ValueError(f"{path}: seed {row['seed_id']} does not have three hits")
            out["seed_id"].append(int(row["seed_id"]))
            out["bottom"].append(hits[0])
            out["middle"].append(hits[1])
            out["top"].append(hits[2])
            for k in ("vertexZ", "quality", "pT", "eta", "phi"):
                out[k].append(np.float32(row[k]))
    res = {k: np.array(v, dtype=np.int64 if k == "seed_id" else np.uint64 if k in ("bottom", "middle", "top") else np.float32)
           for k, v in out.items()}
    if measurement_id is not None:
        lut = {int(v): i for i, v in enumerate(np.asarray(measurement_id))}
        for k in ("bottom", "middle", "top"):
            res[k] = np.array([lut[int(v)] for v in res[k]], dtype=np.uint32)
    # CsvSeedWriter iterates an unordered_map (.cpp:172): restore the container order
    order = np.argsort(res["seed_id"], kind="stable")
    return {k: v[order] for k, v in res.items()}
