"""The C++20 host mirror (acts_b200/host/GridTripletSeedingAlgorithm.hpp) over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "host_mirror_main")


def test_constructor_throws_the_reference_exception_classes(built):
    """std::domain_error / std::runtime_error / std::invalid_argument like the reference ctor chain."""
    res = subprocess.run([BIN, "errors"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "errors ok" in res.stdout


@pytest.mark.gpu
def test_cpp_execute_matches_oracle(built, tmp_path):
    from acts_b200 import config, events
    from oracle import oracle as O

    ev = events.pileup_event(6, mu=20)
    n = ev["x"].size
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(np.uint32(n).tobytes())
        for k in ("x", "y", "z", "r", "varZ", "varR"):
            f.write(np.ascontiguousarray(ev[k], dtype=np.float32).tobytes())
    res = subprocess.run([BIN, "run", str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(fout, "rb").read()
    ns = int(np.frombuffer(raw[:8], np.uint64)[0])
    body = np.frombuffer(raw[8:], np.uint32)
    got = {"bottom": body[:ns], "middle": body[ns:2 * ns], "top": body[2 * ns:3 * ns],
           "quality": body[3 * ns:4 * ns], "vertexZ": body[4 * ns:5 * ns]}
    ref = O.Oracle(config.pu200_config(O.config_init)).run(ev)
    assert ns == ref["quality"].size
    for k in ("bottom", "middle", "top", "quality", "vertexZ"):
        assert np.array_equal(got[k], ref[k].view(np.uint32)), k


def _write_event(path, ev):
    with open(path, "wb") as f:
        f.write(np.uint32(ev["x"].size).tobytes())
        for k in ("x", "y", "z", "r", "varZ", "varR"):
            f.write(np.ascontiguousarray(ev[k], dtype=np.float32).tobytes())


def _read_seed_blocks(path, n_blocks):
    raw = open(path, "rb").read()
    out, o = [], 0
    for _ in range(n_blocks):
        ns = int(np.frombuffer(raw[o:o + 8], np.uint64)[0])
        body = np.frombuffer(raw[o + 8:o + 8 + 20 * ns], np.uint32)
        out.append({"bottom": body[:ns], "middle": body[ns:2 * ns], "top": body[2 * ns:3 * ns],
                    "quality": body[3 * ns:4 * ns], "vertexZ": body[4 * ns:5 * ns]})
        o += 8 + 20 * ns
    return out


@pytest.mark.gpu
def test_cpp_execute_is_reentrant(built, tmp_path):
    """Sequencer.cpp:472-525: execute() is const and entered from several worker threads at once.  Eight threads
    through ONE algorithm object (engine-slot pool inside), every result identical and equal to the oracle's."""
    from acts_b200 import config, events
    from oracle import oracle as O

    evs = [events.pileup_event(20 + i, mu=mu) for i, mu in enumerate((20, 35, 10, 50, 25))]
    files = []
    for i, ev in enumerate(evs):
        files.append(str(tmp_path / f"in{i}.bin"))
        _write_event(files[-1], ev)
    fout = tmp_path / "out.bin"
    res = subprocess.run([BIN, "mt", str(fout), "8", *files], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "mismatches 0" in res.stdout
    assert "slots 4" in res.stdout  # Config::maxConcurrentEvents default: 8 threads share 4 engine slots
    orc = O.Oracle(config.pu200_config(O.config_init))
    for ev, got in zip(evs, _read_seed_blocks(fout, len(evs))):
        ref = orc.run(ev)
        for k in ("bottom", "middle", "top", "quality", "vertexZ"):
            assert np.array_equal(got[k], ref[k].view(np.uint32)), k


@pytest.mark.gpu
def test_cpp_execute_with_vertices(built, tmp_path):
    """Config::inputVertices / vertexZNSigma / vertexZMargin through the mirror class."""
    from acts_b200 import config, events
    from oracle import oracle as O

    ev = events.pileup_event(31, mu=30)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    _write_event(fin, ev)
    vz, vv = [-31.5, 4.25, 60.0], [0.04, 0.25, 1.0]
    args = [str(v) for pair in zip(vz, vv) for v in pair]
    res = subprocess.run([BIN, "runv", str(fin), str(fout), "2.5", "0.5", *args], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    got = _read_seed_blocks(fout, 1)[0]
    cfg = config.pu200_config(O.config_init).update(useVertexZCuts=1, vertexZNSigma=2.5, vertexZMargin=0.5)
    ref = O.Oracle(cfg).run(ev, vertices=(vz, vv))
    assert ref["quality"].size > 0
    for k in ("bottom", "middle", "top", "quality", "vertexZ"):
        assert np.array_equal(got[k], ref[k].view(np.uint32)), k


@pytest.mark.gpu
def test_cpp_execute_strips(built, tmp_path):
    """The strip triplet path through the mirror class (executeStrips -> b200seed_run_strips)."""
    from acts_b200 import config, events
    from oracle import oracle as O

    ev = dict(events.pileup_event(33, mu=20))
    ev["strip"] = events.strip_details(ev, seed=33)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    _write_event(fin, ev)
    with open(fin, "ab") as f:
        f.write(np.ascontiguousarray(ev["strip"], dtype=np.float32).tobytes())
    res = subprocess.run([BIN, "strips", str(fin), str(fout), "0.3"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    got = _read_seed_blocks(fout, 1)[0]
    ref = O.Oracle(config.pu200_config(O.config_init)).run(ev, strip_cot_theta_diff_max=0.3)
    assert ref["quality"].size > 0
    for k in ("bottom", "middle", "top", "quality", "vertexZ"):
        assert np.array_equal(got[k], ref[k].view(np.uint32)), k


@pytest.mark.gpu
def test_cpp_orthogonal_mirror(built, tmp_path):
    """ActsB200::OrthogonalTripletSeedingAlgorithm (acts_b200/host/OrthogonalTripletSeedingAlgorithm.hpp): one object,
    four threads, results identical across threads and equal to the oracle (which is pinned to the reference)."""
    from acts_b200 import config, events
    from oracle import oracle as O

    evs = [events.pileup_event(60 + i, mu=mu) for i, mu in enumerate((15, 30, 8))]
    files = []
    for i, ev in enumerate(evs):
        files.append(str(tmp_path / f"in{i}.bin"))
        _write_event(files[-1], ev)
    fout = tmp_path / "out.bin"
    res = subprocess.run([BIN, "orth", str(fout), "4", *files], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "mismatches 0" in res.stdout
    orc = O.Oracle(*config.orthogonal_config(O.orthogonal_config_init))
    for ev, got in zip(evs, _read_seed_blocks(fout, len(evs))):
        ref = orc.run(ev)
        assert ref["bottom"].size > 0
        for k in ("bottom", "middle", "top", "quality", "vertexZ"):
            assert np.array_equal(got[k], ref[k].view(np.uint32)), k
