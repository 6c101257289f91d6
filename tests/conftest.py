import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    from acts_b200 import build

    build.build_plugin()
    build.build_model()
    build.build_host_mirror_test()
    from oracle import oracle as O

    O.build()
    return True


CONFIGS = ("seeding_py", "pu200", "itk_like", "itk_conf")


def make_config(name, init):
    from acts_b200 import config

    return {"seeding_py": config.seeding_py_config, "pu200": config.pu200_config,
            "itk_like": config.itk_like_config, "itk_conf": config.itk_conf_config}[name](init)
