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
