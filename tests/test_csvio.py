"""ACTS Examples CSV layouts: host-side parsing only, no GPU."""
import numpy as np

from acts_b200 import csvio, events


def test_per_event_filepath_matches_reference_naming():
    # Examples/Framework/src/Utilities/Paths.cpp:42-54
    assert csvio.per_event_filepath("d", "spacepoint.csv", 7) == "d/event000000007-spacepoint.csv"
    assert csvio.per_event_filepath("", "seed.csv", 123456789) == "event123456789-seed.csv"


def test_spacepoint_writer_layout_round_trips_exactly(tmp_path):
    ev = events.pileup_event(3, mu=2)
    path = str(tmp_path / "event000000003-spacepoint.csv")
    csvio.write_spacepoints(path, ev)
    header = open(path).readline().strip()
    assert header == "measurement_id_1,measurement_id_2,geometry_id_1,geometry_id_2,x,y,z,t,var_r,var_z"  # CsvOutputData.hpp:385-393
    sp = csvio.read_spacepoints(path)
    for k in ("x", "y", "z", "varZ", "varR"):
        assert np.array_equal(sp[k].view(np.uint32), ev[k].view(np.uint32)), k
    # r = fastHypot(x, y) (SpacePointMaker.cpp:72) from the float32 coordinates: within one ulp of the generator's
    assert np.allclose(sp["r"], ev["r"], rtol=2e-7, atol=0)
    assert np.array_equal(sp["measurement_id"], np.arange(ev["x"].size, dtype=np.uint64))


def test_spacepoint_reader_layout_is_understood(tmp_path):
    # CsvOutputData.hpp:347-383 (CsvSpacePointReader.cpp:57-75); strip columns are ignored for pixel seeding
    path = str(tmp_path / "event000000000-spacepoints_pixel.csv")
    with open(path, "w") as fh:
        fh.write("measurement_id,sp_x,sp_y,sp_z,sp_radius,sp_covr,sp_covz,sp_topHalfStripLength\n")
        fh.write("5,30.5,-4.25,100.125,30.794682,0.01,0.02,0\n")
        fh.write("9,-70,12,-3,71.021126,0.03,0.04,0\n")
    sp = csvio.read_spacepoints(path)
    assert sp["x"].tolist() == [30.5, -70.0] and sp["z"].tolist() == [100.125, -3.0]
    assert np.allclose(sp["r"], [30.794682, 71.021126])
    assert np.allclose(sp["varR"], [0.01, 0.03]) and np.allclose(sp["varZ"], [0.02, 0.04])
    assert sp["measurement_id"].tolist() == [5, 9]


def test_seed_layout_round_trip(tmp_path):
    ev = events.pileup_event(1, mu=1)
    n = ev["x"].size
    rng = np.random.default_rng(1)
    seeds = {"bottom": rng.integers(0, n, 50).astype(np.uint32), "middle": rng.integers(0, n, 50).astype(np.uint32),
             "top": rng.integers(0, n, 50).astype(np.uint32), "quality": rng.normal(size=50).astype(np.float32),
             "vertexZ": rng.normal(scale=50, size=50).astype(np.float32)}
    ids = np.arange(n, dtype=np.uint64) * 7 + 3
    path = str(tmp_path / "event000000001-seed.csv")
    csvio.write_seeds(path, seeds, ev, measurement_id=ids)
    lines = open(path).read().splitlines()
    # CsvSeedWriter.cpp:168-171
    assert lines[0] == "seed_id,particleId,pT,eta,phi,bX,bY,bZ,mX,mY,mZ,tX,tY,tZ,good/duplicate/fake,vertexZ,quality,Hits_ID"
    assert lines[1].endswith(',"[%d,%d,%d,]"' % (ids[seeds["bottom"][0]], ids[seeds["middle"][0]], ids[seeds["top"][0]]))
    back = csvio.read_seeds(path, measurement_id=ids)
    for k in ("bottom", "middle", "top"):
        assert np.array_equal(back[k], seeds[k])
    for k in ("quality", "vertexZ"):
        assert np.array_equal(back[k].view(np.uint32), seeds[k].view(np.uint32))
