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


def test_strip_columns_round_trip_like_the_reference_reader(tmp_path):
    """The strip columns of the reader layout (CsvOutputData.hpp:354-373) -> the StripCalibrationDetails column the
    way CsvSpacePointReader.cpp:80-110 fills it: half length x unit direction (double) narrowed to float."""
    from acts_b200 import csvio, events

    ev = events.itk_strip_event(0, mu=3)
    path = tmp_path / "event000000000-spacepoint_strip.csv"
    csvio.write_strip_spacepoints(str(path), ev)
    back = csvio.read_spacepoints(str(path))
    n = ev["x"].size
    assert back["x"].size == n and back["strip"].shape == (n, 12) and back["strip"].dtype == np.float32
    for k in ("x", "y", "z", "r", "varZ", "varR"):
        assert np.array_equal(back[k].view(np.uint32), ev[k].view(np.uint32)), k
    # centre and separation are exact; the half vectors go through (length, unit direction): a few ulp
    assert np.array_equal(back["strip"][:, :6].view(np.uint32), ev["strip"][:, :6].view(np.uint32))
    assert np.allclose(back["strip"][:, 6:], ev["strip"][:, 6:], rtol=3e-7, atol=1e-9)
    # a hand-made row: direction (0, 0.6, 0.8), half length 10 -> half vector (0, 6, 8)
    with open(path, "w") as fh:
        fh.write(",".join(csvio.STRIP_READER_COLUMNS) + "\n")
        fh.write("7,1,2,3,2.236068,0.01,0.02,10,5,0,0.6,0.8,1,0,0,0.5,0.25,0.125,100,200,300\n")
    one = csvio.read_spacepoints(str(path))
    want = np.array([[100, 200, 300, 0.5, 0.25, 0.125, 0, 6, 8, 5, 0, 0]], np.float32)
    assert np.allclose(one["strip"], want, rtol=1e-7)
    assert one["measurement_id"][0] == 7
