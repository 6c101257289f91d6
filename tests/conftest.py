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


CONFIGS = ("seeding_py", "pu200", "itk_like", "itk_conf", "itk_pixel", "itk_pixel_grid", "itk_pixel_ho")


def make_config(name, init):
    from acts_b200 import config

    return {"seeding_py": config.seeding_py_config, "pu200": config.pu200_config,
            "itk_like": config.itk_like_config, "itk_conf": config.itk_conf_config,
            # the verbatim ITk PIXEL configuration (itk.py:302-560): with the z-neighbour tables, as
            # addGridTripletSeeding forwards it (without them), and the highOccupancyConfig branch
            "itk_pixel": config.itk_pixel_config,
            "itk_pixel_grid": lambda i: config.itk_pixel_config(i, z_neighbors=False),
            "itk_pixel_ho": lambda i: config.itk_pixel_config(i, high_occupancy=True),
            # the verbatim ITk STRIP configuration (itk.py:458-506: collectors of 100, seedConfirmation)
            "itk_strip": config.itk_strip_config,
            "itk_strip_grid": lambda i: config.itk_strip_config(i, z_neighbors=False)}[name](init)
